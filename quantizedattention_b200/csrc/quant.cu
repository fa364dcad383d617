// HBM-bound pre-pass kernels of the int8 path:
//   * per-(b,h) token mean of K            (K-smoothing, attention_int8.py:24-25 with LEDGER I-1)
//   * per-block int8 quantisation of Q / smoothed K / V / dO
//       scale = fp16(amax|blk| / 127),  value = trunc(fp16(x / scale))   (attention_int8.py:178-195, 241-247)
// One pass over HBM per tensor: 2 B/elem read + 1 B/elem written (+ 2 B/elem for the K mean pass).
// Bit-exact contract: IEEE arithmetic only (no fast-math, no FTZ); x/scale is produced by a
// Markstein-corrected reciprocal multiply, which is correctly rounded because the fp16 divisor
// never has an all-ones fp32 significand.
#include "qa_ptx.cuh"
#include "qa_host.h"

namespace qa {

// ------------------------------------------------------------------------------------------
// K token mean, phase 1: partial sums over a chunk of tokens.  grid = (chunks, B*H), 256 threads.
// Thread layout: D/8 threads across the head dim (16-byte loads), 256/(D/8) token rows per pass.
// ------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) kmean_partial_kernel(const __half* __restrict__ k, float* __restrict__ part,
                                                            int S, int chunk) {
  constexpr int TX = D / 8;          // threads across D
  constexpr int TY = 256 / TX;       // token rows per pass
  const int bh = blockIdx.y, c = blockIdx.x;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int s0 = c * chunk, s1 = min(S, s0 + chunk);
  const uint4* base = reinterpret_cast<const uint4*>(k + (size_t)bh * S * D);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int s = s0 + ty; s < s1; s += TY) {
    uint4 v = __ldg(base + (size_t)s * TX + tx);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      acc[2 * i] += f.x;
      acc[2 * i + 1] += f.y;
    }
  }
  __shared__ float red[TY][D + 1];
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < D) {
    float s = 0.f;
#pragma unroll 4
    for (int r = 0; r < TY; ++r) s += red[r][threadIdx.x];     // fixed order: deterministic
    part[((size_t)bh * gridDim.x + c) * D + threadIdx.x] = s;
  }
}

// phase 2: sum partials in fixed order, divide by S, round to fp16.  One thread per (bh, d).
__global__ void kmean_final_kernel(const float* __restrict__ part, __half* __restrict__ mean, float* __restrict__ sum_out,
                                   int nchunk, int D, int S, int total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int bh = i / D, d = i % D;
  float s = 0.f;
  for (int c = 0; c < nchunk; ++c) s += part[((size_t)bh * nchunk + c) * D + d];
  if (mean != nullptr) mean[i] = __float2half_rn(__fdiv_rn(s, (float)S));
  if (sum_out != nullptr) sum_out[i] = s;       // sequence-sharded ring: token sums are all-reduced before the divide
}

// ------------------------------------------------------------------------------------------
// Block quantisation.  One CTA (256 threads) per quantisation block of `blk` rows x D columns,
// which is contiguous in memory (blk*D fp16).  Each thread keeps ITERS 16-byte vectors in
// registers between the amax pass and the quantise pass, so HBM is read exactly once.
// mean != nullptr: subtract the per-head fp16 mean first (one fp16 rounding), i.e. K-smoothing.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int quant_one(float x, float s, float y, bool nearest) {
  // correctly rounded x / s (y = RN(1/s)), rounded to fp16, then truncated toward zero (the reference, LEDGER I-3)
  // or rounded half-to-even (opt-in accuracy mode)
  float q0 = x * y;
  float r = __fmaf_rn(-s, q0, x);
  float q1 = __fmaf_rn(r, y, q0);
  const float qh = __half2float(__float2half_rn(q1));
  return nearest ? __float2int_rn(qh) : (int)qh;
}

// fp8 (e4m3) variant of the same rule (the reference's "SageAttention3" headline it does not ship,
// README.md:48-54): scale = fp16(amax / 448), value = e4m3(RN(fp16(x / scale))), saturating.  Two values per PTX cvt.
__device__ __forceinline__ uint32_t quant_e4m3x2(float x0, float x1, float s, float y) {
  float q0 = x0 * y, q1 = x1 * y;
  q0 = __fmaf_rn(__fmaf_rn(-s, q0, x0), y, q0);
  q1 = __fmaf_rn(__fmaf_rn(-s, q1, x1), y, q1);
  q0 = __half2float(__float2half_rn(q0));
  q1 = __half2float(__float2half_rn(q1));
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(r) : "f"(q0), "f"(q1));   // first source operand -> upper byte
  return r;
}

template <int ITERS>
__global__ void __launch_bounds__(256) quant_block_kernel(const __half* __restrict__ x, const __half* __restrict__ mean,
                                                          int8_t* __restrict__ out, __half* __restrict__ scales,
                                                          int vec_per_block, int D, int rows_per_head, int blk, int rounding) {
  const size_t b = blockIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(x) + b * vec_per_block;
  uint2* dst = reinterpret_cast<uint2*>(out) + b * vec_per_block;
  const int dvec = D / 8;
  const __half* mrow = nullptr;
  if (mean != nullptr) mrow = mean + (size_t)((b * blk) / rows_per_head) * D;

  uint4 v[ITERS];
#pragma unroll
  for (int i = 0; i < ITERS; ++i) v[i] = __ldcs(src + threadIdx.x + i * 256);   // streaming: read once

  __half2 amax2 = __float2half2_rn(0.f);
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    __half2* h = reinterpret_cast<__half2*>(&v[i]);
    if (mrow != nullptr) {
      const int dv = (threadIdx.x + i * 256) % dvec;
      uint4 mv = __ldg(reinterpret_cast<const uint4*>(mrow) + dv);
      const __half2* mh = reinterpret_cast<const __half2*>(&mv);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __hsub2(h[j], mh[j]);                  // fp16 k - mean (RN)
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) amax2 = __hmax2(amax2, __habs2(h[j]));
  }
  __half amax = __hmax(__low2half(amax2), __high2half(amax2));
  float am = __half2float(amax);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
  __shared__ float wmax[8];
  if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = am;
  __syncthreads();
  am = wmax[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) am = fmaxf(am, wmax[w]);

  const __half s_h = __float2half_rn(__fdiv_rn(am, rounding == 2 ? 448.f : 127.f));     // fp16 amax / 127 (e4m3: / 448), RN
  const float s = __half2float(s_h);
  if (threadIdx.x == 0) scales[b] = s_h;
  const bool zero = (s == 0.f);                                 // LEDGER I-4: all-zero block -> 0
  const float y = zero ? 0.f : __frcp_rn(s);

#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
    if (rounding == 2) {                                        // e4m3
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h[j]);
        w[j] = zero ? 0u : quant_e4m3x2(f.x, f.y, s, y);
      }
      uint2 o;
      o.x = w[0] | (w[1] << 16);
      o.y = w[2] | (w[3] << 16);
      __stcs(dst + threadIdx.x + i * 256, o);
      continue;
    }
    int q[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f = __half22float2(h[j]);
      q[2 * j] = zero ? 0 : quant_one(f.x, s, y, rounding != 0);
      q[2 * j + 1] = zero ? 0 : quant_one(f.y, s, y, rounding != 0);
    }
    uint2 o;
    o.x = (q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
    o.y = (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
    __stcs(dst + threadIdx.x + i * 256, o);
  }
}

}  // namespace qa

using namespace qa;

static int k_mean_impl(const void* k_fp16, void* mean_fp16, void* sum_f32, void* workspace, size_t ws_bytes, int B, int H,
                       int S, int D, void* stream) {
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_k_mean: D must be 64 or 128");
  const int BH = B * H;
  const int chunk = 512;
  const int nchunk = (S + chunk - 1) / chunk;
  const size_t need = (size_t)BH * nchunk * D * sizeof(float);
  if (ws_bytes < need || workspace == nullptr) return qa_fail(QA_ERR_WORKSPACE, "qa_k_mean: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(nchunk, BH);
  if (D == 64)
    kmean_partial_kernel<64><<<grid, 256, 0, st>>>((const __half*)k_fp16, (float*)workspace, S, chunk);
  else
    kmean_partial_kernel<128><<<grid, 256, 0, st>>>((const __half*)k_fp16, (float*)workspace, S, chunk);
  const int total = BH * D;
  kmean_final_kernel<<<(total + 255) / 256, 256, 0, st>>>((const float*)workspace, (__half*)mean_fp16, (float*)sum_f32, nchunk,
                                                         D, S, total);
  return qa_check_launch("qa_k_mean");
}

extern "C" int qa_k_mean(const void* k_fp16, void* mean_fp16, void* workspace, size_t ws_bytes, int B, int H, int S,
                         int D, void* stream) {
  return k_mean_impl(k_fp16, mean_fp16, nullptr, workspace, ws_bytes, B, H, S, D, stream);
}

// fp32 token sums [B*H, D] of a sequence shard (ring KV: all-reduce the sums, then divide by the global length)
extern "C" int qa_k_token_sum(const void* k_fp16, void* sum_f32, void* workspace, size_t ws_bytes, int B, int H, int S,
                              int D, void* stream) {
  return k_mean_impl(k_fp16, nullptr, sum_f32, workspace, ws_bytes, B, H, S, D, stream);
}

extern "C" size_t qa_k_mean_workspace_bytes(int B, int H, int S, int D) {
  return (size_t)B * H * ((S + 511) / 512) * D * sizeof(float);
}

// x: [n_rows, D] fp16 (n_rows = B*H*S); blk rows per quantisation block; mean: [n_rows/rows_per_head, D] fp16 or NULL.
// rounding: 0 = toward zero (the reference), 1 = nearest even (accuracy mode).
extern "C" int qa_quant_block(const void* x_fp16, const void* mean_fp16, void* out_i8, void* scales_fp16, long long n_rows,
                              int D, int blk, int rows_per_head, int rounding, void* stream) {
  if (rounding < 0 || rounding > 2) return qa_fail(QA_ERR_SHAPE, "qa_quant_block: rounding must be 0 (toward zero), 1 (nearest) or 2 (fp8 e4m3)");
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_quant_block: D must be 64 or 128");
  if (blk != 32 && blk != 64 && blk != 128 && blk != 256) return qa_fail(QA_ERR_SHAPE, "qa_quant_block: blk must be 32/64/128/256");
  if (n_rows % blk != 0 || (mean_fp16 && rows_per_head % blk != 0))
    return qa_fail(QA_ERR_SHAPE, "qa_quant_block: rows must be a multiple of the block size");
  if (((uintptr_t)x_fp16 | (uintptr_t)out_i8) & 15) return qa_fail(QA_ERR_ALIGN, "qa_quant_block: 16-byte alignment required");
  const int vec = blk * D / 8;             // 16-byte vectors per block
  const int iters = vec / 256;             // 1..16
  const long long nblk = n_rows / blk;
  cudaStream_t st = (cudaStream_t)stream;
#define QA_LAUNCH_Q(IT)                                                                                              \
  quant_block_kernel<IT><<<(unsigned)nblk, 256, 0, st>>>((const __half*)x_fp16, (const __half*)mean_fp16, (int8_t*)out_i8, \
                                                         (__half*)scales_fp16, vec, D, rows_per_head, blk, rounding)
  switch (iters) {
    case 1: QA_LAUNCH_Q(1); break;
    case 2: QA_LAUNCH_Q(2); break;
    case 4: QA_LAUNCH_Q(4); break;
    case 8: QA_LAUNCH_Q(8); break;
    case 16: QA_LAUNCH_Q(16); break;
    default: return qa_fail(QA_ERR_SHAPE, "qa_quant_block: unsupported block volume");
  }
#undef QA_LAUNCH_Q
  return qa_check_launch("qa_quant_block");
}
