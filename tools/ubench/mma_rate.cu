// tcgen05.mma rate microbenchmark for sm_100a (development tool): clocks per 128 x N x K instruction for kind::i8 and kind::f16
// with K-major and MN-major (transposed) shared-memory operands, SS and TS mode.  One CTA per SM, operands = whatever is in
// shared memory (the values do not matter for the rate).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I quantizedattention_b200/csrc -o tools/ubench/mma_rate tools/ubench/mma_rate.cu -lcuda
#include "qa_ptx.cuh"
#include <cstdio>
using namespace qa;

// mode: 0 i8 SS K/K, 1 i8 SS A MN-major, 2 i8 SS B MN-major, 3 i8 SS both MN-major, 4 f16 SS K/K, 5 f16 SS both MN, 6 i8 TS (B K-major),
//       7 i8 TS (B MN-major), 8 f16 TS (B MN-major)
__global__ void __launch_bounds__(128, 1) mma_rate(long long* out, int mode, int n_mma, int N, int a_col, int d_alt) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done;
  __shared__ uint32_t tbase_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (tid == 0) { mbar_init(&done, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tbase_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tbase_s;
  if (tid == 0) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 32768;
    const bool f16 = (mode == 4 || mode == 5 || mode == 8);
    const uint32_t am = (mode == 1 || mode == 3 || mode == 5) ? 1 : 0, bm = (mode == 2 || mode == 3 || mode == 5 || mode == 7 || mode == 8) ? 1 : 0;
    const uint32_t idesc = f16 ? umma_idesc(1, 0, 0, am, bm, 128, N) : umma_idesc(2, 1, 1, am, bm, 128, N);
    const uint64_t da = am ? umma_smem_desc(a, 16384, 1024, kSwz128) : umma_smem_desc(a, 16, 1024, kSwz128);
    const uint64_t db = bm ? umma_smem_desc(b, 16384, 1024, kSwz128) : umma_smem_desc(b, 16, 1024, kSwz128);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tbase + (d_alt ? (i & 1) * 128 : 0);
      if (mode <= 3) umma_i8_ss(d, da, db, idesc, i > 1);
      else if (mode <= 5) umma_f16_ss(d, da, db, idesc, i > 1);
      else if (mode <= 7) umma_i8_ts(d, tbase + a_col, db, idesc, i > 1);
      else umma_f16_ts(d, tbase + a_col, db, idesc, i > 1);
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
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 66560);
  const char* names[] = {"i8  SS  A K-major  B K-major ", "i8  SS  A MN-major B K-major ", "i8  SS  A K-major  B MN-major", "i8  SS  A MN-major B MN-major",
                         "f16 SS  A K-major  B K-major ", "f16 SS  A MN-major B MN-major", "i8  TS  A TMEM     B K-major ", "i8  TS  A TMEM     B MN-major",
                         "f16 TS  A TMEM     B MN-major"};
  for (int ctas : {1, 148}) {
    printf("%d CTA(s): clocks per instruction (128 x 128 x 32 for i8, 128 x 128 x 16 for f16), issue / complete\n", ctas);
    for (int m = 0; m < 9; ++m) {
      long long h[2];
      for (int rep = 0; rep < 2; ++rep) {
        mma_rate<<<ctas, 128, 66560>>>(d, m, 256, 128, 256, 1);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("  %s : %7.1f / %7.1f %s\n", names[m], h[0] / 256.0, h[1] / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  printf("TS / SS variants, 1 CTA: mode, N, A column, D alternating -> clocks per instruction\n");
  for (int m : {0, 6, 4, 8})
    for (int N : {64, 128, 256})
      for (int ac : {256, 384, 448})
        for (int da : {0, 1}) {
          if (N == 256 && (da == 1 || ac < 384)) continue;      // D = 256 columns at 0
          long long h[2];
          for (int rep = 0; rep < 2; ++rep) { mma_rate<<<1, 128, 66560>>>(d, m, 256, N, ac, da); cudaDeviceSynchronize(); }
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("  %s N=%3d a_col=%3d d_alt=%d : %7.1f\n", names[m], N, ac, da, h[1] / 256.0);
        }
  return 0;
}
