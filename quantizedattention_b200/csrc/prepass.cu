// HBM-bound backward / JVP pre- and post-passes.
//   delta = rowsum(dO * O)            (fp32 pre-pass replacing the reference's per-tile recompute,
//                                      attention_int8.py:397-398, attention_bf16.py:416; LEDGER I-12 / B-8)
//   fp32 -> bf16 / fp16 conversions   (backward dO operand, JVP operands; LEDGER J-2)
//   fp32 -> fp16 cast of the dQ accumulation workspace
#include "qa_ptx.cuh"
#include "qa_host.h"

namespace qa {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

// one warp per row; D in {64, 128}; optional bf16 copy of dO
template <typename TdO, typename TO>
__global__ void __launch_bounds__(256) delta_kernel(const TdO* __restrict__ dO, const TO* __restrict__ O,
                                                    float* __restrict__ delta, __nv_bfloat16* __restrict__ dO_bf16,
                                                    long long n_rows, int D) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  float s = 0.f;
  for (int d = lane * 2; d < D; d += 64) {
    const float a0 = to_f(dO[row * D + d]), a1 = to_f(dO[row * D + d + 1]);
    const float b0 = to_f(O[row * D + d]), b1 = to_f(O[row * D + d + 1]);
    s = fmaf(a0, b0, s);
    s = fmaf(a1, b1, s);
    if (dO_bf16 != nullptr)
      *reinterpret_cast<__nv_bfloat162*>(dO_bf16 + row * D + d) = __floats2bfloat162_rn(a0, a1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) delta[row] = s;
}

// elementwise fp32 -> {fp16, bf16}; n multiple of 4
template <typename Tout>
__global__ void __launch_bounds__(256) cast_f32_kernel(const float4* __restrict__ in, Tout* __restrict__ out, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 v = __ldcs(in + i);
  if constexpr (sizeof(Tout) == 2 && std::is_same<Tout, __half>::value) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(out + i * 4) = o;
  } else {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(out + i * 4) = o;
  }
}

}  // namespace qa

using namespace qa;

// in_dtype: 0 = fp16 dO/O (int8 path), 1 = fp32 dO/O (bf16 path; also writes a bf16 copy of dO if dO_bf16 != NULL)
extern "C" int qa_bwd_delta(const void* dO, const void* O, void* delta_f32, void* dO_bf16, long long n_rows, int D,
                            int in_dtype, void* stream) {
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_bwd_delta: D must be 64 or 128");
  const unsigned grid = (unsigned)((n_rows + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == 0)
    delta_kernel<__half, __half><<<grid, 256, 0, st>>>((const __half*)dO, (const __half*)O, (float*)delta_f32, nullptr, n_rows, D);
  else
    delta_kernel<float, float><<<grid, 256, 0, st>>>((const float*)dO, (const float*)O, (float*)delta_f32,
                                                    (__nv_bfloat16*)dO_bf16, n_rows, D);
  return qa_check_launch("qa_bwd_delta");
}

// out_dtype: 0 = fp16, 1 = bf16
extern "C" int qa_cast_f32(const void* in_f32, void* out, long long n, int out_dtype, void* stream) {
  if (n % 4) return qa_fail(QA_ERR_SHAPE, "qa_cast_f32: element count must be a multiple of 4");
  if (((uintptr_t)in_f32 & 15) || ((uintptr_t)out & 7)) return qa_fail(QA_ERR_ALIGN, "qa_cast_f32: alignment");
  const long long n4 = n / 4;
  const unsigned grid = (unsigned)((n4 + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == 0) cast_f32_kernel<__half><<<grid, 256, 0, st>>>((const float4*)in_f32, (__half*)out, n4);
  else cast_f32_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const float4*)in_f32, (__nv_bfloat16*)out, n4);
  return qa_check_launch("qa_cast_f32");
}

// ---------------------------------------------------------------------------------------------------------
// dQ epilogue of the int8 backward: dq = fp16(dq_ws + sm_scale * rowsum[row] * k_mean[head][col]).
// The second term is the K-smoothing correction (LEDGER I-1); rowsum / k_mean may be null (no smoothing).
// ---------------------------------------------------------------------------------------------------------
namespace qa {
template <int D>
__global__ void __launch_bounds__(256) dq_finalize_kernel(const float4* __restrict__ ws, const float* __restrict__ rowsum,
                                                          const __half* __restrict__ k_mean, __half* __restrict__ out,
                                                          long long n_rows, int S, float sm_scale) {
  constexpr int V = D / 4;                                          // float4 per row
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * V) return;
  const long long row = i / V;
  const int c = (int)(i % V) * 4;
  float4 v = ws[i];
  if (rowsum != nullptr) {
    const float rs = rowsum[row] * sm_scale;
    const __half2* km = reinterpret_cast<const __half2*>(k_mean + (row / S) * D + c);
    const float2 k0 = __half22float2(km[0]), k1 = __half22float2(km[1]);
    v.x = fmaf(rs, k0.x, v.x); v.y = fmaf(rs, k0.y, v.y); v.z = fmaf(rs, k1.x, v.z); v.w = fmaf(rs, k1.y, v.w);
  }
  __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(out + i * 4) = o;
}
}  // namespace qa

extern "C" int qa_int8_bwd_finalize(const void* dq_ws_f32, const void* rowsum_ws_f32, const void* k_mean_f16, void* dq_f16,
                                    int BH, int S, int D, void* stream) {
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd_finalize: D must be 64 or 128");
  if ((rowsum_ws_f32 == nullptr) != (k_mean_f16 == nullptr))
    return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd_finalize: rowsum workspace and k_mean go together");
  if (((uintptr_t)dq_ws_f32 & 15) || ((uintptr_t)dq_f16 & 7)) return qa_fail(QA_ERR_ALIGN, "qa_int8_bwd_finalize: alignment");
  const long long n_rows = (long long)BH * S;
  const long long n = n_rows * (D / 4);
  const unsigned grid = (unsigned)((n + 255) / 256);
  const float sm_scale = (float)(1.0 / sqrt((double)D));
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 128)
    qa::dq_finalize_kernel<128><<<grid, 256, 0, st>>>((const float4*)dq_ws_f32, (const float*)rowsum_ws_f32,
                                                      (const __half*)k_mean_f16, (__half*)dq_f16, n_rows, S, sm_scale);
  else
    qa::dq_finalize_kernel<64><<<grid, 256, 0, st>>>((const float4*)dq_ws_f32, (const float*)rowsum_ws_f32,
                                                     (const __half*)k_mean_f16, (__half*)dq_f16, n_rows, S, sm_scale);
  return qa_check_launch("qa_int8_bwd_finalize");
}
