// tcgen05.cp (shared memory -> TMEM, 32x128b.warpx4: one 512-byte scale-factor atom) rate microbenchmark for sm_100a (development
// tool): clocks per instruction, alone and interleaved with block-scaled MMAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I quantizedattention_b200/csrc -o tools/ubench/cp_rate tools/ubench/cp_rate.cu -lcuda
#include "qa_ptx.cuh"
#include <cstdio>
using namespace qa;

__global__ void __launch_bounds__(128, 1) cp_rate(long long* out, int mode, int n) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;
  if (tid == 0) { mbar_init(&done, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tbase_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tbase_s;
  if (tid == 0) {
    const uint32_t a = smem_u32(smem);
    const uint32_t idesc = umma_idesc_bs(1, 1, 0, 0, 128, 128, 0);
    const uint64_t da = umma_smem_desc(a + 4096, 16, 512, kSwz64), db = umma_smem_desc(a + 16384, 16, 512, kSwz64);
    const uint64_t sd = umma_smem_desc(a, 0, 128, kSwzNone);
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      if (mode == 0 || mode == 2)
        asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tbase + 384 + 4 * (i & 7)), "l"(sd) : "memory");
      if (mode == 1 || mode == 2) umma_nvf4_ss(tbase + (i & 1) * 128, da, db, idesc, tbase + 384, tbase + 392, 1);
    }
    long long t1 = clock64();
    umma_commit(&done);
    mbar_wait(&done, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(cp_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  const char* names[] = {"tcgen05.cp alone", "nvf4 SS MMA alone", "cp + MMA alternating (per pair)"};
  for (int m = 0; m < 3; ++m) {
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) { cp_rate<<<1, 128, 40960>>>(d, m, 256); cudaDeviceSynchronize(); }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("%-34s: issue %7.1f  complete %7.1f clk per iteration %s\n", names[m], h[0] / 256.0, h[1] / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
