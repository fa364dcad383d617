// NVFP4 (microscaling) quantisation pre-passes of the fp4 forward (the SageAttention3 headline the
// reference names, /root/reference/README.md:48-54, and does not ship).  HBM-bound: 2 B/elem read, 0.5 B/elem + 1/16 B/elem of
// scales written, + 2 B/elem for the per-head amax pass of Q / K.  For V (and optionally Q / K) the per-head amax the two-level
// scale needs is produced inside the same kernel: every CTA keeps its tile in shared memory / registers, publishes the tile's amax with atomicMax, counts itself in and waits until all CTAs of its head have
// done so (CTAs are dispatched in blockIdx order and a head's CTAs are contiguous, so the earliest incomplete head is always
// fully resident: the wait cannot deadlock), then quantises from the data it already holds.
//
// Two-level scaling (the NVFP4 recipe): per head  sg = amax_head / (6 * 448)  (fp32), per block of 16 elements along the
// CONTRACTION axis  sf = e4m3_rn(amax_blk / 6 / sg)  in [0, 448],  value = e2m1_rn(x / (sf * sg))  in [-6, 6]; all
// arithmetic IEEE fp32 (no fast-math; quant_e2m1x8 reproduces the IEEE quotient's code without dividing), so the codes and scales are bit-exact against oracle/fp4_ref.py.
//   Q, K : blocks along D (contraction of Q K^T); codes [B*H*S, D/2] bytes, element 2i in the low nibble of byte i.
//   V    : blocks along the KEY axis (contraction of P V); codes stored transposed [B*H, D, S/2] because tcgen05
//          kind::mxf4nvf4 takes 4-bit operands K-major only.
// Scale factors are written in the layout tcgen05.cp.32x128b.warpx4 expects (profiles/r02_fp4_probe.txt): per 128-row
// tile and per 64-element K step one 512-byte atom, byte 16 * (r % 32) + 4 * (r / 32) + s = scale of row r, block s.
#include "qa_ptx.cuh"
#include "qa_host.h"
#include <cuda_fp4.h>
#include <cuda_fp8.h>

namespace qa {

__device__ __forceinline__ float e4m3_to_float(uint8_t c) {
  const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)c, __NV_E4M3);
  return __half2float(*reinterpret_cast<const __half*>(&h));
}
__device__ __forceinline__ uint8_t float_to_e4m3(float x) { return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3); }
__device__ __forceinline__ uint32_t pack_e2m1x8(const float (&y)[8]) {
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w |= (uint32_t)__nv_cvt_float2_to_fp4x2(make_float2(y[2 * i], y[2 * i + 1]), __NV_E2M1, cudaRoundNearest) << (8 * i);
  return w;
}

// e2m1_rn(fl(v / scale)) for 8 values, without the 8 IEEE divisions: the quotient is bracketed by v * r (1 -+ 2^-20) with r an
// approximate reciprocal (error of r, of the two products and of the IEEE quotient itself together < 2^-21 relative, so the
// IEEE quotient lies strictly inside the bracket) and the conversion is monotonic: when both ends give the same eight codes
// the IEEE quotient gives them too.  Otherwise (a rounding threshold inside a 2^-19-wide window: ~1e-5 of the elements) the
// block is redone with the divisions.  The result is bit-identical to the division form in every case.
__device__ __forceinline__ uint32_t quant_e2m1x8(const float* v, float scale) {
  if (!(scale > 0.f)) return 0u;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(scale));
  const float rl = __fmul_rn(r, 1.0f - 0x1p-20f), rh = __fmul_rn(r, 1.0f + 0x1p-20f);
  float a[8], b[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { a[e] = __fmul_rn(v[e], rl); b[e] = __fmul_rn(v[e], rh); }
  const uint32_t lo = pack_e2m1x8(a), hi = pack_e2m1x8(b);
  if (lo == hi && scale >= 0x1p-100f) return lo;                  // (a tiny scale would overflow the reciprocal)
  float y[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) y[e] = __fdiv_rn(v[e], scale);
  return pack_e2m1x8(y);
}

// Two-pass fallback (a head with more 128-row tiles than the GPU has SMs cannot be guaranteed resident, so its CTAs must not
// wait for each other): amax over a head of |x - mean| (fp16 rounding of the difference; abs and max are exact in fp16, so the
// reduction stays in half2).  grid = (chunks, BH); the grid stride is a multiple of D / 8, so a thread keeps its 8 columns.
template <int D>
__global__ void __launch_bounds__(256) fp4_head_amax_kernel(const __half* __restrict__ x, const __half* __restrict__ mean,
                                                            float* __restrict__ amax, int S, int S_valid) {
  constexpr unsigned DV = D / 8;
  const int bh = blockIdx.y;
  const unsigned n8 = (unsigned)S_valid * DV;                    // padding rows of a ragged sequence count as zeros
  const uint4* base = reinterpret_cast<const uint4*>(x + (size_t)bh * S * D);
  const unsigned i0 = blockIdx.x * 256u + threadIdx.x;
  uint4 mu = make_uint4(0u, 0u, 0u, 0u);                          // x - (+0) = x
  if (mean != nullptr) mu = __ldg(reinterpret_cast<const uint4*>(mean + (size_t)bh * D) + (i0 % DV));
  const __half2* mu2 = reinterpret_cast<const __half2*>(&mu);
  __half2 m2 = __float2half2_rn(0.f);
  for (unsigned i = i0; i < n8; i += gridDim.x * 256u) {
    const uint4 v = __ldg(base + i);
    const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) m2 = __hmax2(m2, __habs2(__hsub2(h2[e], mu2[e])));
  }
  float m = fmaxf(__low2float(m2), __high2float(m2));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    atomicMax(reinterpret_cast<int*>(amax + bh), __float_as_int(m));       // non-negative floats order like ints
  }
}

// Head-wide amax inside one kernel: CTA-reduce `m`, atomicMax into amax[bh] (non-negative floats order like ints), count this
// CTA in and spin until all `ncta` CTAs of the head have arrived.  Returns the head's amax to every thread.
__device__ __forceinline__ float fp4_head_amax_sync(float m, float* amax, unsigned* count, int bh, unsigned ncta, float* red) {
  if (count == nullptr) return amax[bh];                         // two-pass mode: fp4_head_amax_kernel ran before
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) red[w] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < nw; ++i) m = fmaxf(m, red[i]);
    atomicMax(reinterpret_cast<int*>(amax + bh), __float_as_int(m));
    __threadfence();
    atomicAdd(count + bh, 1u);
    while (atomicAdd(count + bh, 0u) < ncta) __nanosleep(64);
    __threadfence();
    red[0] = __int_as_float(atomicMax(reinterpret_cast<int*>(amax + bh), 0));     // coherent read of the final value
  }
  __syncthreads();
  return red[0];
}

// Q / K: one thread per 16-element block along D.  CTA = ROWS rows x (D/16) blocks = ROWS / 128 of a scale-factor tile (small
// CTAs: a 1024-thread CTA holds its slots from its first load to its last store and only two fit on an SM).
template <int D, int ROWS>
__global__ void __launch_bounds__(ROWS * D / 16) fp4_quant_rows_kernel(const __half* __restrict__ x, const __half* __restrict__ mean,
                                                                       float* __restrict__ amax, unsigned* __restrict__ count,
                                                                       uint8_t* __restrict__ codes, uint8_t* __restrict__ sf,
                                                                       float* __restrict__ sg_out, int S, int S_valid) {
  constexpr int NB = D / 16;
  constexpr unsigned PER_TILE = 128 / ROWS;
  __shared__ float red[32];
  const unsigned tile = blockIdx.x / PER_TILE, nt = (unsigned)S / 128u;    // 128-row tile over B*H*S: the S / 128 tiles of a head are contiguous
  const int r = (int)(blockIdx.x % PER_TILE) * ROWS + threadIdx.x / NB, b = threadIdx.x % NB;
  const size_t row = (size_t)tile * 128 + r;
  const int bh = (int)(tile / nt);
  const int row_in_head = (int)(tile % nt) * 128 + r;
  const uint4* src = reinterpret_cast<const uint4*>(x + row * D + b * 16);
  const bool valid = row_in_head < S_valid;                      // padding rows stay zero after the smoothing
  float v[16];
  __half2 am2 = __float2half2_rn(0.f);
#pragma unroll
  for (int hv = 0; hv < 2; ++hv) {
    uint4 u = make_uint4(0u, 0u, 0u, 0u), mu = u;
    if (valid) {
      u = __ldg(src + hv);
      if (mean != nullptr) mu = __ldg(reinterpret_cast<const uint4*>(mean + (size_t)bh * D + b * 16) + hv);
    }
    const __half2 *h2 = reinterpret_cast<const __half2*>(&u), *mu2 = reinterpret_cast<const __half2*>(&mu);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 t = __hsub2(h2[e], mu2[e]);                   // (x - (+0) = x without a mean)
      am2 = __hmax2(am2, __habs2(t));                             // abs and max are exact in fp16
      const float2 f = __half22float2(t);
      v[hv * 8 + 2 * e] = f.x; v[hv * 8 + 2 * e + 1] = f.y;
    }
  }
  const float am = fmaxf(__low2float(am2), __high2float(am2));
  const float sg = __fdiv_rn(fp4_head_amax_sync(am, amax, count, bh, (unsigned)(S / ROWS), red), 2688.0f);
  if (sg_out != nullptr && row_in_head == 0 && b == 0) sg_out[bh] = sg;
  uint8_t sc = 0;
  float scale = 0.f;
  if (sg > 0.f) {
    sc = float_to_e4m3(__fdiv_rn(__fdiv_rn(am, 6.0f), sg));
    scale = __fmul_rn(e4m3_to_float(sc), sg);
  }
  uint32_t w[2];
#pragma unroll
  for (int hv = 0; hv < 2; ++hv) w[hv] = quant_e2m1x8(v + hv * 8, scale);
  *reinterpret_cast<uint2*>(codes + row * (D / 2) + b * 8) = make_uint2(w[0], w[1]);
  // scale-factor atoms of this tile: K step = b / 4, block s = b % 4
  sf[(size_t)tile * (NB / 4) * 512 + (b / 4) * 512 + 16 * (r % 32) + 4 * (r / 32) + (b % 4)] = sc;
}

// V: CTA = one 128-key tile of one head, transposed through shared memory; thread = (d, block of 16 keys)
template <int D>
__global__ void __launch_bounds__(256) fp4_quant_vt_kernel(const __half* __restrict__ v, float* __restrict__ amax,
                                                           unsigned* __restrict__ count, uint8_t* __restrict__ codes_t,
                                                           uint8_t* __restrict__ sf, float* __restrict__ sg_out, int S) {
  __shared__ __half tile[128][D + 8];                            // +8 halves: 16-byte row skew against bank conflicts
  __shared__ float red[32];
  __shared__ uint2 out_codes[D][9];                              // the tile's codes [d][64 bytes] (+8: bank skew) and its two scale-factor
  __shared__ __align__(16) uint8_t out_sf[1024];                 // atoms, staged so that they leave in 16-byte pieces
  const int nt = S / 128;
  const int bh = blockIdx.x / nt, j = blockIdx.x % nt;           // 1-D grid, a head's tiles contiguous (see fp4_head_amax_sync)
  const uint4* src = reinterpret_cast<const uint4*>(v + ((size_t)bh * S + (size_t)j * 128) * D);
  float tm = 0.f;
  for (int i = threadIdx.x; i < 128 * D / 8; i += 256) {
    const int rr = i / (D / 8), c8 = i % (D / 8);
    const uint4 u = __ldg(src + i);
    *reinterpret_cast<uint4*>(&tile[rr][c8 * 8]) = u;
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(h2[e]); tm = fmaxf(tm, fmaxf(fabsf(f.x), fabsf(f.y))); }
  }
  const float sg = __fdiv_rn(fp4_head_amax_sync(tm, amax, count, bh, (unsigned)nt, red), 2688.0f);   // (its barriers also publish the tile)
  if (sg_out != nullptr && j == 0 && threadIdx.x == 0) sg_out[bh] = sg;
  for (int it = threadIdx.x; it < D * 8; it += 256) {
    const int b = it / D, d = it % D;                             // consecutive threads read consecutive d of one key row: no bank conflicts
    float x[16];
    float am = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) { x[e] = __half2float(tile[b * 16 + e][d]); am = fmaxf(am, fabsf(x[e])); }
    uint8_t sc = 0;
    float scale = 0.f;
    if (sg > 0.f) {
      sc = float_to_e4m3(__fdiv_rn(__fdiv_rn(am, 6.0f), sg));
      scale = __fmul_rn(e4m3_to_float(sc), sg);
    }
    uint32_t w[2];
#pragma unroll
    for (int hv = 0; hv < 2; ++hv) w[hv] = quant_e2m1x8(x + hv * 8, scale);
    out_codes[d][b] = make_uint2(w[0], w[1]);
    // rows of the B operand of P V are the D output columns; K step = b / 4 (64 keys), block s = b % 4
    out_sf[(b / 4) * 512 + 16 * (d % 32) + 4 * (d / 32) + (b % 4)] = sc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D * 4; i += 256) {                // 64 bytes per output column d, 16 per thread
    const int d = i / 4, c = i % 4;
    const uint2 lo = out_codes[d][2 * c], hi = out_codes[d][2 * c + 1];
    *reinterpret_cast<uint4*>(codes_t + ((size_t)bh * D + d) * (S / 2) + (size_t)j * 64 + c * 16) = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
  if (threadIdx.x < 64)
    reinterpret_cast<uint4*>(sf + ((size_t)bh * (S / 128) + j) * 1024)[threadIdx.x] = reinterpret_cast<const uint4*>(out_sf)[threadIdx.x];
}

}  // namespace qa

using namespace qa;

static int fp4_sm_count() {                                       // SMs of the current device (the residency bound of the one-pass mode)
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return n;
}

// x: [BH, S, D] fp16; mean: [BH, D] fp16 or NULL (K smoothing); amax_ws: 2 * BH 32-bit words of scratch (overwritten);
// codes: [BH*S, D/2] bytes; sf: [BH*S/128][D/64][512] bytes; sg: [BH] fp32.  D = 128, S % 128 == 0.
extern "C" int qa_fp4_quant_rows_ragged(const void* x_fp16, const void* mean_fp16, void* amax_ws, void* codes, void* sf, void* sg_f32,
                                        int BH, int S, int S_valid, int D, void* stream) {
  if (S_valid <= 0 || S_valid > S) return qa_fail(QA_ERR_SHAPE, "qa_fp4_quant_rows: S_valid must lie in (0, S]");
  if (D != 128) return qa_fail(QA_ERR_SHAPE, "qa_fp4_quant_rows: D must be 128");
  if (BH <= 0 || S <= 0 || S % 128) return qa_fail(QA_ERR_SHAPE, "qa_fp4_quant_rows: S must be a positive multiple of 128");
  if (!x_fp16 || !amax_ws || !codes || !sf || !sg_f32) return qa_fail(QA_ERR_ALIGN, "qa_fp4_quant_rows: null pointer");
  if (((uintptr_t)x_fp16 | (uintptr_t)codes | (uintptr_t)sf) & 15) return qa_fail(QA_ERR_ALIGN, "qa_fp4_quant_rows: 16-byte alignment required");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(amax_ws, 0, (size_t)BH * 8, st);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  // Two passes for Q / K: CTAs that wait for the rest of their head leave too few loads in flight (measured at B*H = 64, S = 8192:
  // 0.129 ms in one pass vs 0.078 ms in two with 256-thread CTAs; 0.18 vs 0.125 ms with 1024-thread CTAs); the V kernel (tile in
  // shared memory) gains from the single pass (0.105 -> 0.099 ms)
  const bool one_pass = false;
  if (!one_pass) {
    const int chunks = (int)(((size_t)S * D / 8 + 256 * 16 - 1) / (256 * 16));
    fp4_head_amax_kernel<128><<<dim3(chunks, BH), 256, 0, st>>>((const __half*)x_fp16, (const __half*)mean_fp16, (float*)amax_ws, S, S_valid);
  }
  fp4_quant_rows_kernel<128, 32><<<(unsigned)((size_t)BH * S / 32), 256, 0, st>>>((const __half*)x_fp16, (const __half*)mean_fp16,
                                                                                  (float*)amax_ws, one_pass ? (unsigned*)amax_ws + BH : nullptr,
                                                                                  (uint8_t*)codes, (uint8_t*)sf, (float*)sg_f32, S, S_valid);
  return qa_check_launch("qa_fp4_quant_rows");
}

extern "C" int qa_fp4_quant_rows(const void* x_fp16, const void* mean_fp16, void* amax_ws, void* codes, void* sf, void* sg_f32, int BH,
                                 int S, int D, void* stream) {
  return qa_fp4_quant_rows_ragged(x_fp16, mean_fp16, amax_ws, codes, sf, sg_f32, BH, S, S, D, stream);
}

// v: [BH, S, D] fp16 -> codes_t: [BH, D, S/2] bytes (transposed), sf: [BH*S/128][2][512] bytes, sg: [BH] fp32; amax_ws: 2 * BH words
extern "C" int qa_fp4_quant_vt(const void* v_fp16, void* amax_ws, void* codes_t, void* sf, void* sg_f32, int BH, int S, int D,
                               void* stream) {
  if (D != 128) return qa_fail(QA_ERR_SHAPE, "qa_fp4_quant_vt: D must be 128");
  if (BH <= 0 || S <= 0 || S % 128) return qa_fail(QA_ERR_SHAPE, "qa_fp4_quant_vt: S must be a positive multiple of 128");
  if (!v_fp16 || !amax_ws || !codes_t || !sf || !sg_f32) return qa_fail(QA_ERR_ALIGN, "qa_fp4_quant_vt: null pointer");
  if (((uintptr_t)v_fp16 | (uintptr_t)codes_t | (uintptr_t)sf) & 15) return qa_fail(QA_ERR_ALIGN, "qa_fp4_quant_vt: 16-byte alignment required");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(amax_ws, 0, (size_t)BH * 8, st);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  const bool one_pass = S / 128 <= fp4_sm_count();
  if (!one_pass) {
    const int chunks = (int)(((size_t)S * D / 8 + 256 * 16 - 1) / (256 * 16));
    fp4_head_amax_kernel<128><<<dim3(chunks, BH), 256, 0, st>>>((const __half*)v_fp16, nullptr, (float*)amax_ws, S, S);
  }
  fp4_quant_vt_kernel<128><<<(unsigned)((size_t)BH * (S / 128)), 256, 0, st>>>((const __half*)v_fp16, (float*)amax_ws,
                                                                             one_pass ? (unsigned*)amax_ws + BH : nullptr, (uint8_t*)codes_t,
                                                                             (uint8_t*)sf, (float*)sg_f32, S);
  return qa_check_launch("qa_fp4_quant_vt");
}
