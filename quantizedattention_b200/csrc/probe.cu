// Self-checking hardware probes (debug entry points of the C-ABI library, used by tests/test_probe_gpu.py).
// They pin down, on the real B200, the layout/descriptor conventions the attention kernels rely on:
//   qa_probe_mma : one CTA stages host-built shared-memory images, issues tcgen05.mma with caller-chosen
//                  descriptors (SS or TS mode) and dumps the TMEM accumulator.
//   qa_probe_tma : one CTA performs a TMA tiled load with a given swizzle and dumps shared memory.
#include "qa_ptx.cuh"
#include "qa_host.h"

namespace qa {

struct ProbeMmaParams {
  const uint8_t* a_img;   // SS: smem image of A; TS: uint32 [128][a_tmem_cols] TMEM image
  const uint8_t* b_img;   // smem image of B
  uint32_t* d_out;        // [128][n_cols] accumulator dump (raw 32-bit)
  int a_bytes, b_bytes;
  int a_lbo, a_sbo, a_layout, a_kstep_bytes;
  int b_lbo, b_sbo, b_layout, b_kstep_bytes;
  uint32_t idesc;
  int kind;               // 0 = f16, 1 = i8, 2 = tf32
  int n_mma, n_cols;
  int a_in_tmem, a_tmem_cols, a_tmem_kstep_cols;
};

__global__ void __launch_bounds__(128) probe_mma_kernel(ProbeMmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((p.a_in_tmem ? 0 : p.a_bytes) + 1023) / 1024 * 1024;

  if (!p.a_in_tmem)
    for (int i = tid * 16; i < p.a_bytes; i += 128 * 16)
      *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(p.a_img + i);
  for (int i = tid * 16; i < p.b_bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(p.b_img + i);
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t a_col0 = 256;   // TS-mode A operand lives at columns [256, 256 + a_tmem_cols)

  if (p.a_in_tmem) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(p.a_img) + (size_t)tid * p.a_tmem_cols;
    for (int c = 0; c < p.a_tmem_cols; c += 8) {
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = row[c + i];
      tmem_st8(lane_addr + a_col0 + c, r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (tid == 0) {
    for (int k = 0; k < p.n_mma; ++k) {
      uint64_t bd = umma_smem_desc(smem_u32(sb) + k * p.b_kstep_bytes, p.b_lbo, p.b_sbo, p.b_layout);
      uint32_t acc = k > 0;
      if (p.a_in_tmem) {
        uint32_t at = tbase + a_col0 + k * p.a_tmem_kstep_cols;
        if (p.kind == 1) umma_i8_ts(tbase, at, bd, p.idesc, acc);
        else umma_f16_ts(tbase, at, bd, p.idesc, acc);
      } else {
        uint64_t ad = umma_smem_desc(smem_u32(sa) + k * p.a_kstep_bytes, p.a_lbo, p.a_sbo, p.a_layout);
        if (p.kind == 1) umma_i8_ss(tbase, ad, bd, p.idesc, acc);
        else if (p.kind == 0) umma_f16_ss(tbase, ad, bd, p.idesc, acc);
        else umma_tf32_ss(tbase, ad, bd, p.idesc, acc);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < p.n_cols; c += 8) {
    uint32_t r[8];
    tmem_ld8(lane_addr + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) p.d_out[(size_t)tid * p.n_cols + c + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

__global__ void __launch_bounds__(128) probe_tma_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t* out, int bytes,
                                                        int rank, int c0, int c1, int c2) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < bytes; i += 128) smem[i] = 0xEE;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, bytes);
    if (rank == 2) tma_load_2d(smem, &tmap, &bar, c0, c1);
    else tma_load_3d(smem, &tmap, &bar, c0, c1, c2);
  }
  mbar_wait(&bar, 0);
  for (int i = tid; i < bytes; i += 128) out[i] = smem[i];
}


// ---------------------------------------------------------------------------------------------------------
// Block-scaled MMA probe (NVFP4 / MXFP4 microscaling): stages host-built images of A, B (packed e2m1,
// two per byte) and of the scale factors, copies the scale factors shared memory -> TMEM with tcgen05.cp.32x128b.warpx4
// (one 512-byte atom = 32 rows x 16 bytes -> 4 TMEM columns, replicated over the four lane quadrants), issues
// tcgen05.mma.kind::mxf4nvf4.block_scale and dumps the fp32 accumulator.
//   sf atom (512 B):  byte 16 * (r % 32) + 4 * (r / 32) + s  =  scale of row r, K block s of this MMA (K = 64: s < 4)
// ---------------------------------------------------------------------------------------------------------
struct ProbeMmaBsParams {
  const uint8_t *a_img, *b_img, *sfa_img, *sfb_img;
  uint32_t* d_out;
  int a_bytes, b_bytes, sfa_bytes, sfb_bytes;
  int a_lbo, a_sbo, a_layout, a_kstep_bytes;
  int b_lbo, b_sbo, b_layout, b_kstep_bytes;
  uint32_t idesc;
  int kind;               // 0 = mxf4nvf4 block16 (ue4m3 scales, K = 64), 1 = mxf4 block32 (ue8m0, K = 64), 2 = mxf8f6f4 block32 (K = 32)
  int n_mma, n_cols;
  int sfa_cols_per_mma, sfb_cols_per_mma;   // TMEM columns the scale factors advance per MMA; bits 8.. of sfb_cols_per_mma: first column offset of B's scales
  int a_in_tmem, a_tmem_cols, a_tmem_kstep_cols;
};

__device__ __forceinline__ void tmem_cp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

__global__ void __launch_bounds__(128) probe_mma_bs_kernel(ProbeMmaBsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  auto up = [](int v) { return (v + 1023) / 1024 * 1024; };
  uint8_t* sa = smem;
  uint8_t* sb = sa + up(p.a_in_tmem ? 0 : p.a_bytes);
  uint8_t* sfa = sb + up(p.b_bytes);
  uint8_t* sfb = sfa + up(p.sfa_bytes);
  auto stage = [&](uint8_t* dst, const uint8_t* src, int bytes) {
    for (int i = tid * 16; i < bytes; i += 128 * 16) *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
  };
  if (!p.a_in_tmem) stage(sa, p.a_img, p.a_bytes);
  stage(sb, p.b_img, p.b_bytes);
  stage(sfa, p.sfa_img, p.sfa_bytes);
  stage(sfb, p.sfb_img, p.sfb_bytes);
  fence_proxy_async_smem();
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t a_col0 = 256, sfa_col0 = 384, sfb_col0 = 448;

  if (p.a_in_tmem) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(p.a_img) + (size_t)tid * p.a_tmem_cols;
    for (int c = 0; c < p.a_tmem_cols; c += 8) {
      uint32_t r[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = row[c + i];
      tmem_st8(lane_addr + a_col0 + c, r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  if (tid == 0) {
    // scale factors: 512-byte atoms (32 rows x 16 B, 8-row core matrices 128 B apart), one tcgen05.cp per atom -> 4 columns
    for (int i = 0; i < p.sfa_bytes / 512; ++i)
      tmem_cp_32x128b_warpx4(tbase + sfa_col0 + 4 * i, umma_smem_desc(smem_u32(sfa) + 512 * i, 0, 128, kSwzNone));
    for (int i = 0; i < p.sfb_bytes / 512; ++i)
      tmem_cp_32x128b_warpx4(tbase + sfb_col0 + 4 * i, umma_smem_desc(smem_u32(sfb) + 512 * i, 0, 128, kSwzNone));
    for (int k = 0; k < p.n_mma; ++k) {                       // tcgen05.cp and tcgen05.mma execute in issue order
      const uint64_t bd = umma_smem_desc(smem_u32(sb) + k * p.b_kstep_bytes, p.b_lbo, p.b_sbo, p.b_layout);
      const uint32_t acc = k > 0, tsfa = tbase + sfa_col0 + k * p.sfa_cols_per_mma, tsfb = tbase + sfb_col0 + (p.sfb_cols_per_mma >> 8) + k * (p.sfb_cols_per_mma & 0xff);
      if (p.a_in_tmem) {
        const uint32_t at = tbase + a_col0 + k * p.a_tmem_kstep_cols;
        umma_nvf4_ts(tbase, at, bd, p.idesc, tsfa, tsfb, acc);
      } else {
        const uint64_t ad = umma_smem_desc(smem_u32(sa) + k * p.a_kstep_bytes, p.a_lbo, p.a_sbo, p.a_layout);
        if (p.kind == 0) umma_nvf4_ss(tbase, ad, bd, p.idesc, tsfa, tsfb, acc);
        else if (p.kind == 1) umma_mxf4_ss(tbase, ad, bd, p.idesc, tsfa, tsfb, acc);
        else umma_mxf8_ss(tbase, ad, bd, p.idesc, tsfa, tsfb, acc);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < p.n_cols; c += 8) {
    uint32_t r[8];
    tmem_ld8(lane_addr + c, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) p.d_out[(size_t)tid * p.n_cols + c + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

}  // namespace qa

using namespace qa;

extern "C" int qa_probe_mma(const void* a_img, int a_bytes, const void* b_img, int b_bytes, void* d_out, int a_lbo,
                            int a_sbo, int a_layout, int a_kstep_bytes, int b_lbo, int b_sbo, int b_layout,
                            int b_kstep_bytes, unsigned idesc, int kind, int n_mma, int n_cols, int a_in_tmem,
                            int a_tmem_cols, int a_tmem_kstep_cols, void* stream) {
  ProbeMmaParams p;
  p.a_img = (const uint8_t*)a_img; p.b_img = (const uint8_t*)b_img; p.d_out = (uint32_t*)d_out;
  p.a_bytes = a_bytes; p.b_bytes = b_bytes;
  p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.a_layout = a_layout; p.a_kstep_bytes = a_kstep_bytes;
  p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.b_layout = b_layout; p.b_kstep_bytes = b_kstep_bytes;
  p.idesc = idesc; p.kind = kind; p.n_mma = n_mma; p.n_cols = n_cols;
  p.a_in_tmem = a_in_tmem; p.a_tmem_cols = a_tmem_cols; p.a_tmem_kstep_cols = a_tmem_kstep_cols;
  if (n_cols % 8 || n_cols > 256 || (a_bytes & 15) || (b_bytes & 15)) return qa_fail(QA_ERR_SHAPE, "qa_probe_mma: bad sizes");
  size_t smem = 1024 + ((size_t)(a_in_tmem ? 0 : a_bytes) + 1023) / 1024 * 1024 + b_bytes;
  if (smem > 200 * 1024) return qa_fail(QA_ERR_SHAPE, "qa_probe_mma: images too large");
  cudaFuncSetAttribute(probe_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_mma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  return qa_check_launch("qa_probe_mma");
}

extern "C" int qa_probe_mma_bs(const void* a_img, int a_bytes, const void* b_img, int b_bytes, const void* sfa_img, int sfa_bytes,
                               const void* sfb_img, int sfb_bytes, void* d_out, int a_lbo, int a_sbo, int a_layout,
                               int a_kstep_bytes, int b_lbo, int b_sbo, int b_layout, int b_kstep_bytes, unsigned idesc, int kind,
                               int n_mma, int n_cols, int sfa_cols_per_mma, int sfb_cols_per_mma, int a_in_tmem, int a_tmem_cols,
                               int a_tmem_kstep_cols, void* stream) {
  ProbeMmaBsParams p;
  p.a_img = (const uint8_t*)a_img; p.b_img = (const uint8_t*)b_img; p.sfa_img = (const uint8_t*)sfa_img; p.sfb_img = (const uint8_t*)sfb_img;
  p.d_out = (uint32_t*)d_out;
  p.a_bytes = a_bytes; p.b_bytes = b_bytes; p.sfa_bytes = sfa_bytes; p.sfb_bytes = sfb_bytes;
  p.a_lbo = a_lbo; p.a_sbo = a_sbo; p.a_layout = a_layout; p.a_kstep_bytes = a_kstep_bytes;
  p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.b_layout = b_layout; p.b_kstep_bytes = b_kstep_bytes;
  p.idesc = idesc; p.kind = kind; p.n_mma = n_mma; p.n_cols = n_cols;
  p.sfa_cols_per_mma = sfa_cols_per_mma; p.sfb_cols_per_mma = sfb_cols_per_mma;
  p.a_in_tmem = a_in_tmem; p.a_tmem_cols = a_tmem_cols; p.a_tmem_kstep_cols = a_tmem_kstep_cols;
  if (n_cols % 8 || n_cols > 256 || (a_bytes & 15) || (b_bytes & 15) || (sfa_bytes % 512) || (sfb_bytes % 512) || sfa_bytes > 8192 ||
      sfb_bytes > 8192)
    return qa_fail(QA_ERR_SHAPE, "qa_probe_mma_bs: bad sizes");
  auto up = [](size_t v) { return (v + 1023) / 1024 * 1024; };
  size_t smem = 1024 + up(a_in_tmem ? 0 : a_bytes) + up(b_bytes) + up(sfa_bytes) + up(sfb_bytes);
  if (smem > 200 * 1024) return qa_fail(QA_ERR_SHAPE, "qa_probe_mma_bs: images too large");
  cudaFuncSetAttribute(probe_mma_bs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_mma_bs_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  return qa_check_launch("qa_probe_mma_bs");
}

// elem_bytes in {1,2,4}; dims/box innermost-first; strides in bytes for dims >= 1.
extern "C" int qa_probe_tma(const void* gptr, int elem_bytes, int rank, const unsigned long long* dims,
                            const unsigned long long* strides_bytes, const unsigned* box, int swizzle, const int* coords,
                            void* out, void* stream) {
  CUtensorMap tm;
  CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                           : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                             : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  uint64_t d[3], s[3];
  uint32_t b[3];
  size_t bytes = elem_bytes;
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; bytes *= box[i]; if (i) s[i - 1] = strides_bytes[i - 1]; }
  int rc = qa_make_tmap(&tm, gptr, dt, rank, d, s, b, swizzle);
  if (rc) return rc;
  size_t smem = bytes + 1024;
  cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_tma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tm, (uint8_t*)out, (int)bytes, rank, coords[0], coords[1],
                                                           rank > 2 ? coords[2] : 0);
  return qa_check_launch("qa_probe_tma");
}

// ---------------------------------------------------------------------------------------------------------
// TMEM -> register read bandwidth probe: every warp streams tcgen05.ld over its lane quadrant (4 KiB per load, 32
// registers per thread) in one of the instruction shapes, with `DEPTH` loads in flight before the wait.
// Gives the measured ceiling for kernels whose accumulators must be drained from TMEM every tile.
// ---------------------------------------------------------------------------------------------------------
namespace qa {
#define QA_LD32(NAME, MNEMONIC)                                                                                        \
  __device__ __forceinline__ void NAME(uint32_t taddr, uint32_t (&r)[32]) {                                            \
    asm volatile("tcgen05.ld.sync.aligned." MNEMONIC ".b32 "                                                           \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26," \
                 "%27,%28,%29,%30,%31}, [%32];"                                                                        \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),            \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),          \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),          \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                               \
                 : "r"(taddr));                                                                                        \
  }
QA_LD32(ld_32x32b, "32x32b.x32")
QA_LD32(ld_16x256b, "16x256b.x8")
QA_LD32(ld_16x128b, "16x128b.x16")
QA_LD32(ld_16x64b, "16x64b.x32")
#undef QA_LD32

template <int SHAPE, int DEPTH>
__global__ void __launch_bounds__(DEPTH == 1 ? 1024 : 512, 1) probe_tmem_bw_kernel(uint32_t* sink, int iters) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t lane_addr = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; c += DEPTH) {
      uint32_t r[DEPTH][32];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        const uint32_t a = lane_addr + ((warp >> 2) * 128 + (c + d) * 64) % 448;   // 16-lane shapes span 64 columns
        if (SHAPE == 0) ld_32x32b(a, r[d]);
        else if (SHAPE == 1) ld_16x256b(a, r[d]);      // 16 lanes x 64 columns
        else if (SHAPE == 2) ld_16x128b(a, r[d]);
        else ld_16x64b(a, r[d]);
      }
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[d][i];
    }
  }
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}
}  // namespace qa

// Launches `blocks` CTAs of `threads` threads; each warp issues iters * 4 loads of 4 KiB.
// shape: 0 = 32x32b.x32, 1 = 16x256b.x8, 2 = 16x128b.x16, 3 = 16x64b.x32; depth: 1 or 2 loads in flight per warp.
extern "C" int qa_probe_tmem_bw_ex(void* sink, int blocks, int threads, int iters, int shape, int depth, void* stream) {
  if (threads % 128 || threads > 1024) return qa_fail(QA_ERR_SHAPE, "qa_probe_tmem_bw: threads must be a multiple of 128");
  if (shape < 0 || shape > 3 || (depth != 1 && depth != 2) || (depth == 2 && threads > 512))
    return qa_fail(QA_ERR_SHAPE, "qa_probe_tmem_bw: bad shape/depth");
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* s = (uint32_t*)sink;
#define QA_GO(S, DP) qa::probe_tmem_bw_kernel<S, DP><<<blocks, threads, 0, st>>>(s, iters)
#define QA_SH(S) (depth == 1 ? QA_GO(S, 1) : QA_GO(S, 2))
  if (shape == 0) QA_SH(0); else if (shape == 1) QA_SH(1); else if (shape == 2) QA_SH(2); else QA_SH(3);
#undef QA_SH
#undef QA_GO
  return qa_check_launch("qa_probe_tmem_bw");
}

extern "C" int qa_probe_tmem_bw(void* sink, int blocks, int threads, int iters, void* stream) {
  return qa_probe_tmem_bw_ex(sink, blocks, threads, iters, 0, 1, stream);
}
