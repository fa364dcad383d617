// Recompute backward of the bf16 path (reference attention_bf16.py:361-444 under the
// 8-LEDGER contract B-4..B-8: dS = P*(dP - delta), sm_scale, zero-weight strict-causal masking, delta pre-pass).
//
// One CTA = one 128-key tile of one (batch, head), looping over query tiles (k-outer, as the reference).
//   S  = Q_i K_j^T   (fp16 x fp16)      dP = dO_i V_j^T (bf16 x bf16)            -> TMEM cols 0..127 / 128..255
//   dV_j += P^T dO_i (bf16)             dK_j += dS^T Q_i (fp16)                  -> fp32 accumulators RESIDENT in TMEM
//   dQ_i += dS K_j   (fp16)             -> TMEM cols 0..127 (aliasing S), drained through red.global.add.f32
// P is rounded to bf16 and dS to fp16 for the tensor cores (fp32 accumulation); the reference computes these in
// fp32 (TF32 tl.dot) -- tolerance-level parity, stated in tests/test_bf16_bwd_gpu.py.
#include "qa_ptx.cuh"
#include "qa_host.h"
#include <type_traits>

namespace qa {

constexpr int kBAtom = 128 * 128;

template <int D, int STAGES>
struct Bf16BwdSmem {
  static constexpr int kTile = 128 * D * 2;
  static constexpr int kPT = 128 * 128 * 2;
  static constexpr int off_k = 0;
  static constexpr int off_v = off_k + kTile;
  static constexpr int off_q = off_v + kTile;
  static constexpr int off_do = off_q + STAGES * kTile;
  static constexpr int off_p = off_do + STAGES * kTile;
  static constexpr int off_ds = off_p + kPT;
  static constexpr int total = off_ds + kPT + 1024;
};

struct Bf16BwdParams {
  const float* lse;      // [BH*S]
  const float* delta;    // [BH*S]
  float *dq, *dk, *dv;   // fp32 [BH*S, D]; dq zero-initialised by the caller
  int S, causal;
  int S_valid;           // rows [S_valid, S) of every head are zero padding (ragged sequence): padded keys get P = 0
  float sm_scale, qk_scale;
};

__device__ __forceinline__ void red_add_v4f(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int D, int STAGES>
__global__ void __launch_bounds__(288, 1)
bf16_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do, Bf16BwdParams p) {
  using L = Bf16BwdSmem<D, STAGES>;
  constexpr int DH = D / 2;
  constexpr int kDAtoms = D / 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t kv_full, qdo_full[2], sd_full, pds_full, parts_full, tmem_free;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, j = blockIdx.x;
  const int nq = p.S / 128;
  const int i0 = p.causal ? j : 0;                   // query tiles before the diagonal see none of these keys
  const size_t head_row0 = (size_t)bh * p.S;

  if (tid == 0) {
    mbar_init(&kv_full, 1); mbar_init(&qdo_full[0], 1); mbar_init(&qdo_full[1], 1);
    mbar_init(&sd_full, 1); mbar_init(&pds_full, 8); mbar_init(&parts_full, 1); mbar_init(&tmem_free, 8);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 8) {
    if (elect_one()) {
      constexpr uint32_t id_s = umma_idesc(1, 0, 0, 0, 0, 128, 128);      // S : f16, K-major x K-major
      constexpr uint32_t id_dp = umma_idesc(1, 1, 1, 0, 0, 128, 128);     // dP: bf16
      constexpr uint32_t id_dv = umma_idesc(1, 1, 1, 1, 1, 128, D);       // dV: bf16, A^T (MN) x B MN
      constexpr uint32_t id_dk = umma_idesc(1, 0, 0, 1, 1, 128, D);       // dK: f16
      constexpr uint32_t id_dq = umma_idesc(1, 0, 0, 0, 1, 128, D);       // dQ: f16, A K-major x B MN
      const uint32_t a_k = smem_u32(smem + L::off_k), a_v = smem_u32(smem + L::off_v);
      const uint32_t a_p = smem_u32(smem + L::off_p), a_ds = smem_u32(smem + L::off_ds);
      auto load_qdo = [&](int i, int st) {
        mbar_expect_tx(&qdo_full[st], 2 * L::kTile);
#pragma unroll
        for (int a = 0; a < kDAtoms; ++a) {
          tma_load_2d(smem + L::off_q + st * L::kTile + a * kBAtom, &tm_q, &qdo_full[st], a * 64, (int)head_row0 + i * 128);
          tma_load_2d(smem + L::off_do + st * L::kTile + a * kBAtom, &tm_do, &qdo_full[st], a * 64, (int)head_row0 + i * 128);
        }
      };
      mbar_expect_tx(&kv_full, 2 * L::kTile);
#pragma unroll
      for (int a = 0; a < kDAtoms; ++a) {
        tma_load_2d(smem + L::off_k + a * kBAtom, &tm_k, &kv_full, a * 64, (int)head_row0 + j * 128);
        tma_load_2d(smem + L::off_v + a * kBAtom, &tm_v, &kv_full, a * 64, (int)head_row0 + j * 128);
      }
      load_qdo(i0, 0);
      mbar_wait(&kv_full, 0);
      for (int i = i0, n = 0; i < nq; ++i, ++n) {
        const int st = (STAGES == 2) ? (n & 1) : 0;
        const uint32_t a_q = smem_u32(smem + L::off_q + st * L::kTile), a_do = smem_u32(smem + L::off_do + st * L::kTile);
        mbar_wait(&qdo_full[st], (STAGES == 2) ? ((n >> 1) & 1) : (n & 1));
        if (n > 0) mbar_wait(&tmem_free, (n - 1) & 1);
        tc_fence_after();
        if (STAGES == 2 && i + 1 < nq) load_qdo(i + 1, st ^ 1);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t o = (k >> 2) * kBAtom + (k & 3) * 32;
          umma_f16_ss(tbase + 0, umma_smem_desc(a_q + o, 16, 1024, kSwz128), umma_smem_desc(a_k + o, 16, 1024, kSwz128), id_s, k > 0);
          umma_f16_ss(tbase + 128, umma_smem_desc(a_do + o, 16, 1024, kSwz128), umma_smem_desc(a_v + o, 16, 1024, kSwz128), id_dp, k > 0);
        }
        umma_commit(&sd_full);
        mbar_wait(&pds_full, n & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {                       // contraction over 128 query rows (dV, dK) / 128 keys (dQ)
          const uint64_t pT = umma_smem_desc(a_p + k * 2048, kBAtom, 1024, kSwz128);
          const uint64_t dsT = umma_smem_desc(a_ds + k * 2048, kBAtom, 1024, kSwz128);
          const uint64_t dsK = umma_smem_desc(a_ds + (k >> 2) * kBAtom + (k & 3) * 32, 16, 1024, kSwz128);
          umma_f16_ss(tbase + 256, pT, umma_smem_desc(a_do + k * 2048, kBAtom, 1024, kSwz128), id_dv, (n > 0) || (k > 0));
          umma_f16_ss(tbase + 384, dsT, umma_smem_desc(a_q + k * 2048, kBAtom, 1024, kSwz128), id_dk, (n > 0) || (k > 0));
          umma_f16_ss(tbase + 0, dsK, umma_smem_desc(a_k + k * 2048, kBAtom, 1024, kSwz128), id_dq, k > 0);
        }
        umma_commit(&parts_full);
        if (STAGES == 1 && i + 1 < nq) {                    // single stage: reload once this tile's MMAs are done
          mbar_wait(&parts_full, n & 1);
          load_qdo(i + 1, 0);
        }
      }
    }
  } else {
    const int half = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    int n = 0;
    for (int i = i0; i < nq; ++i, ++n) {
      const uint32_t ph = n & 1;
      const int qi = i * 128 + row;                          // query index in the head
      const size_t qrow = head_row0 + qi;
      const float lse = p.lse[qrow];
      const float dlt = p.delta[qrow];
      const bool diag = p.causal && (i == j);
      const bool tailk = (j + 1) * 128 > p.S_valid;            // this CTA's key tile is the ragged last one
      mbar_wait(&sd_full, ph);
      tc_fence_after();
      const float2 qk2 = make_float2(p.qk_scale, p.qk_scale), nlse2 = make_float2(-lse, -lse), ndlt2 = make_float2(-dlt, -dlt);
      auto compute = [&](auto masked) {                            // masked = diagonal tile or causal query row 0
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t rs[32], rp[32];
          tmem_ld32(lane_addr + half * 64 + ch * 32, rs);
          tmem_ld32(lane_addr + 128 + half * 64 + ch * 32, rp);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint32_t wp[4], wd[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = g * 8 + e * 2;
              const float2 ex = __ffma2_rn(make_float2(__uint_as_float(rs[c]), __uint_as_float(rs[c + 1])), qk2, nlse2);
              float2 pv = make_float2(ex2_approx(ex.x), ex2_approx(ex.y));                       // :391-392
              if (decltype(masked)::value) {
                const int key = j * 128 + half * 64 + ch * 32 + c;
                if ((diag && key >= qi) || (p.causal && qi == 0) || key >= p.S_valid) pv.x = 0.f;          // strict causal, weight 0; row 0: fixup
                if ((diag && key + 1 >= qi) || (p.causal && qi == 0) || key + 1 >= p.S_valid) pv.y = 0.f;  // ragged: padded keys
              }
              const float2 ds = __fmul2_rn(pv, __fadd2_rn(make_float2(__uint_as_float(rp[c]), __uint_as_float(rp[c + 1])), ndlt2));   // dS = P*(dP - delta)
              __nv_bfloat162 pb = __float22bfloat162_rn(pv);
              __half2 dh = __float22half2_rn(ds);
              wp[e] = *reinterpret_cast<uint32_t*>(&pb);
              wd[e] = *reinterpret_cast<uint32_t*>(&dh);
            }
            const uint32_t off = (uint32_t)half * kBAtom + swz128(row, (ch * 32 + g * 8) * 2);
            sts128(smem_u32(smem) + L::off_p + off, wp[0], wp[1], wp[2], wp[3]);
            sts128(smem_u32(smem) + L::off_ds + off, wd[0], wd[1], wd[2], wd[3]);
          }
        }
      };
      if ((p.causal && (diag || i == 0)) || tailk) compute(std::true_type{}); else compute(std::false_type{});
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full);
      // ---- drain dQ partial
      mbar_wait(&parts_full, ph);
      tc_fence_after();
      float* dq_dst = p.dq + qrow * D + half * DH;
#pragma unroll
      for (int ch = 0; ch < DH / 32; ++ch) {
        uint32_t r[32];
        tmem_ld32(lane_addr + half * DH + ch * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float2 sm2 = make_float2(p.sm_scale, p.sm_scale);
          const float2 a = __fmul2_rn(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), sm2);
          const float2 b = __fmul2_rn(make_float2(__uint_as_float(r[c + 2]), __uint_as_float(r[c + 3])), sm2);
          red_add_v4f(dq_dst + ch * 32 + c, a.x, a.y, b.x, b.y);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_free);
    }
    // ---- epilogue: dV_j, dK_j accumulators from TMEM (row = key)
    tc_fence_after();
    const size_t krow = head_row0 + (size_t)j * 128 + row;
    float* dv_dst = p.dv + krow * D + half * DH;
    float* dk_dst = p.dk + krow * D + half * DH;
#pragma unroll
    for (int ch = 0; ch < DH / 32; ++ch) {
      uint32_t r[32];
      tmem_ld32(lane_addr + 256 + half * DH + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        *reinterpret_cast<float4*>(dv_dst + ch * 32 + c) =
            make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
      tmem_ld32(lane_addr + 384 + half * DH + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        *reinterpret_cast<float4*>(dk_dst + ch * 32 + c) =
            make_float4(__uint_as_float(r[c]) * p.sm_scale, __uint_as_float(r[c + 1]) * p.sm_scale,
                        __uint_as_float(r[c + 2]) * p.sm_scale, __uint_as_float(r[c + 3]) * p.sm_scale);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<512>(tbase);
}

// causal row 0 attends uniformly to ALL keys (LEDGER B-1): dV[k] += dO[0] / S for every key, no dS.
__global__ void bf16_bwd_row0_fixup_kernel(const float* __restrict__ dO, float* __restrict__ dv, int S, int S_valid, int D) {
  const int bh = blockIdx.y;
  const int k = blockIdx.x * (blockDim.x / D) + threadIdx.x / D, d = threadIdx.x % D;
  if (k < S_valid) dv[((size_t)bh * S + k) * D + d] += dO[(size_t)bh * S * D + d] / (float)S_valid;
}

template <int D, int STAGES>
static int launch_bf16_bwd(const void* q, const void* k, const void* v, const void* do_bf16, const float* dO_f32,
                           const Bf16BwdParams& p, int BH, cudaStream_t st) {
  using L = Bf16BwdSmem<D, STAGES>;
  CUtensorMap tq, tk, tv, tdo;
  uint64_t dims[2] = {(uint64_t)D, (uint64_t)BH * p.S};
  uint64_t str[1] = {(uint64_t)D * 2};
  uint32_t box[2] = {64, 128};
  int rc;
  if ((rc = qa_make_tmap(&tq, q, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tk, k, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tv, v, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tdo, do_bf16, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  auto kern = bf16_bwd_kernel<D, STAGES>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  dim3 grid(p.S / 128, BH);
  kern<<<grid, 288, L::total, st>>>(tq, tk, tv, tdo, p);
  int r = qa_check_launch("qa_bf16_bwd");
  if (r) return r;
  if (p.causal) {
    const int kpb = 256 / D;
    dim3 g2((p.S + kpb - 1) / kpb, BH);
    bf16_bwd_row0_fixup_kernel<<<g2, 256, 0, st>>>(dO_f32, p.dv, p.S, p.S_valid, D);
    r = qa_check_launch("qa_bf16_bwd(row0)");
  }
  return r;
}

// attn_bf16_bwd2.cu: warp-specialised kernel, D = 128
int launch_bf16_bwd_ws(const void* q, const void* k, const void* v, const void* do_bf16, const float* dO_f32, const float* lse,
                       const float* delta, float* dq, float* dk, float* dv, int BH, int S, int S_valid, int causal, cudaStream_t st);

}  // namespace qa

using namespace qa;

// q, k fp16; v bf16; dO_bf16 = bf16 copy of dO (from qa_bwd_delta); dO_f32 the original; lse, delta fp32 [BH*S];
// dq (zero-initialised), dk, dv: fp32 [BH*S, D].
// variant 0: default (D = 128: warp-specialised kernel with transposed logits, attn_bf16_bwd2.cu; D = 64: the phase-sequential
// kernel above); variant 1: the phase-sequential kernel for every D.
// Ragged sequences: buffers zero-padded per head to S (a multiple of 128); rows [S_valid, S) are padding: the caller pads dO
// with zeros and lse with a large finite value (P = 0 for padded query rows), padded keys get P = 0 in the kernel.
extern "C" int qa_bf16_bwd_ragged(const void* q_f16, const void* k_f16, const void* v_bf16, const void* dO_bf16, const void* dO_f32,
                                  const void* lse_f32, const void* delta_f32, void* dq_f32, void* dk_f32, void* dv_f32, int BH, int S,
                                  int S_valid, int D, int causal, int variant, void* stream) {
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_bf16_bwd: D must be 64 or 128");
  if (S % 128 || S <= 0 || BH <= 0) return qa_fail(QA_ERR_SHAPE, "qa_bf16_bwd: S must be a positive multiple of 128");
  if (S_valid <= S - 128 || S_valid > S) return qa_fail(QA_ERR_SHAPE, "qa_bf16_bwd: S_valid must lie in (S - 128, S]");
  if (variant < 0 || variant > 1) return qa_fail(QA_ERR_SHAPE, "qa_bf16_bwd: variant must be 0 or 1");
  const void* ptrs[] = {q_f16, k_f16, v_bf16, dO_bf16, lse_f32, delta_f32, dq_f32, dk_f32, dv_f32};
  for (const void* ptr : ptrs)
    if (!ptr || (reinterpret_cast<uintptr_t>(ptr) & 15)) return qa_fail(QA_ERR_ALIGN, "qa_bf16_bwd: null or not 16-byte aligned pointer");
  if (causal && !dO_f32) return qa_fail(QA_ERR_ALIGN, "qa_bf16_bwd: causal needs dO_f32");
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 128 && variant == 0) {
    if (causal && (reinterpret_cast<uintptr_t>(dO_f32) & 15)) return qa_fail(QA_ERR_ALIGN, "qa_bf16_bwd: dO_f32 not 16-byte aligned");
    // causal row 0 (uniform weight on every key, LEDGER B-1) is added in the kernel's dV write-back: no fix-up pass
    return launch_bf16_bwd_ws(q_f16, k_f16, v_bf16, dO_bf16, (const float*)dO_f32, (const float*)lse_f32, (const float*)delta_f32,
                              (float*)dq_f32, (float*)dk_f32, (float*)dv_f32, BH, S, S_valid, causal, st);
  }
  Bf16BwdParams p;
  p.lse = (const float*)lse_f32; p.delta = (const float*)delta_f32;
  p.dq = (float*)dq_f32; p.dk = (float*)dk_f32; p.dv = (float*)dv_f32;
  p.S = S; p.S_valid = S_valid; p.causal = causal;
  p.sm_scale = (float)(1.0 / sqrt((double)D));
  p.qk_scale = (float)((1.0 / sqrt((double)D)) * 1.44269504);
  return D == 128 ? launch_bf16_bwd<128, 1>(q_f16, k_f16, v_bf16, dO_bf16, (const float*)dO_f32, p, BH, st)
                  : launch_bf16_bwd<64, 2>(q_f16, k_f16, v_bf16, dO_bf16, (const float*)dO_f32, p, BH, st);
}

extern "C" int qa_bf16_bwd_ex(const void* q_f16, const void* k_f16, const void* v_bf16, const void* dO_bf16, const void* dO_f32,
                              const void* lse_f32, const void* delta_f32, void* dq_f32, void* dk_f32, void* dv_f32, int BH, int S,
                              int D, int causal, int variant, void* stream) {
  return qa_bf16_bwd_ragged(q_f16, k_f16, v_bf16, dO_bf16, dO_f32, lse_f32, delta_f32, dq_f32, dk_f32, dv_f32, BH, S, S, D, causal,
                            variant, stream);
}

// q, k fp16; v bf16; dO_bf16 = bf16 copy of dO (from qa_bwd_delta); dO_f32 the original; lse, delta fp32 [BH*S];
// dq (zero-initialised), dk, dv: fp32 [BH*S, D].
extern "C" int qa_bf16_bwd(const void* q_f16, const void* k_f16, const void* v_bf16, const void* dO_bf16, const void* dO_f32,
                           const void* lse_f32, const void* delta_f32, void* dq_f32, void* dk_f32, void* dv_f32, int BH, int S,
                           int D, int causal, void* stream) {
  return qa_bf16_bwd_ex(q_f16, k_f16, v_bf16, dO_bf16, dO_f32, lse_f32, delta_f32, dq_f32, dk_f32, dv_f32, BH, S, D, causal, 0, stream);
}
