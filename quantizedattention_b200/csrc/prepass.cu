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
