// Warp-specialised recompute backward of the bf16 path for D = 128 (reference
// attention_bf16.py:361-444 under the 8-LEDGER contract B-4..B-8).  Same arithmetic as bf16_bwd_kernel
// (attn_bf16_bwd.cu): P rounded to bf16 for dV, dS = P*(dP - delta) from the fp32 P rounded to fp16, fp32 accumulation.
//
// One CTA = one 128-key tile j of one (batch, head), looping over the query tiles i (k-outer, as the reference).  The
// logits are computed TRANSPOSED so that the probabilities can feed the tensor core from TMEM:
//   S^T  = K_j Q_i^T        (fp16)  -> TMEM A   [lane = key, column = query]
//   dP^T = V_j dO_i^T       (bf16)  -> TMEM B
//   dV_j += P^T dO_i        (bf16)  P^T read from TMEM (written over S^T by the compute warps)  -> TMEM, resident
//   dK_j += dS^T Q_i        (fp16)  dS^T from shared memory (K-major A)                          -> TMEM, resident
//   dQ_i^T = K_j^T dS^T     (fp16)  both operands MN-major, -> TMEM B [lane = d, column = query]: a drain warp's
//                                   red.global.add then covers 32 consecutive floats of one dQ row (one 128 B line)
// Shared memory: K, V, two stages of Q and dO, dS^T = 7 x 32 KB (P never touches shared memory: that is what makes
// room for the second Q / dO stage at D = 128).
// Roles (16 warps, setmaxnreg): warps 0..3 drain dQ^T, 4..11 compute P^T and dS^T (two per TMEM lane quadrant, 64 query
// columns each), 12 issues tcgen05.mma, 13 issues TMA.  Issue order of the MMA warp per query tile n:
//   dV(n) | S(n+1) | dQ(n), dK(n) | dP(n+1)
// so that the compute warps' dS(n) phase runs under dV(n) / S(n+1), their P(n+1) phase under dQ(n) / dK(n), and the dQ
// drain under dK(n).
#include "qa_ptx.cuh"
#include "qa_host.h"

namespace qa {

struct Bf16BwdParams2 {
  const float* lse;      // [BH*S] log2 domain
  const float* delta;    // [BH*S]
  float *dq, *dk, *dv;   // fp32 [BH*S, D]; dq zero-initialised by the caller
  int S, causal;
  float sm_scale, qk_scale;
};

struct Bf16Bwd2Smem {
  static constexpr int kTile = 128 * 128 * 2;
  static constexpr int off_k = 0;
  static constexpr int off_v = kTile;
  static constexpr int off_q = 2 * kTile;          // 2 stages
  static constexpr int off_do = 4 * kTile;         // 2 stages
  static constexpr int off_ds = 6 * kTile;
  static constexpr int off_ld = 7 * kTile;         // [stage][lse | delta][128] fp32
  static constexpr int off_bar = off_ld + 2048;
  static constexpr int used = off_bar + 160;
  static constexpr int total = 232448;             // everything an SM has; the align-up pad must fit in total - used
};

enum Bf16Bwd2Bar { KV_FULL = 0, Q_FULL0, Q_FULL1, DO_FULL0, DO_FULL1, Q_FREE0, Q_FREE1, DO_FREE0, DO_FREE1, S_FULL, P_READY,
                   DP_FULL, DS_READY, DS_FREE, DQ_FULL, DQ_FREE, ACC_FULL, kNumBars };

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(512, 1)
bf16_bwd_ws_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do, Bf16BwdParams2 p) {
  using L = Bf16Bwd2Smem;
  constexpr int D = 128;
  constexpr int kAtom = 128 * 128;                     // one 64-column (128 B) swizzle atom column of 128 rows
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if ((int)(smem - smem_raw) > L::total - L::used) __trap();
  float* ld_s = reinterpret_cast<float*>(smem + L::off_ld);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::off_bar);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem + L::off_bar + kNumBars * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, j = blockIdx.x;
  const int nq = p.S / 128;
  const int i0 = p.causal ? j : 0;                     // query tiles before the diagonal see none of these keys
  const int nt = nq - i0;
  const size_t head_row0 = (size_t)bh * p.S;

  if (tid == 0) {
    for (int b = 0; b < kNumBars; ++b) mbar_init(&bars[b], (b == P_READY || b == DS_READY) ? 8 : (b == DQ_FREE ? 4 : 1));
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc<512>(tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_base_s;
  constexpr uint32_t tA = 0, tB = 128, tDV = 256, tDK = 384;

  if (warp < 4) {
    // =========================== dQ^T drain: lane = d, column = query ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    const int d = warp * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16) + tB;
    const float sm = p.sm_scale;
    for (int n = 0; n < nt; ++n) {
      float* dst = p.dq + (head_row0 + (size_t)(i0 + n) * 128) * D + d;
      mbar_wait(&bars[DQ_FULL], n & 1);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t r[64];
        tmem_ld64(lane_addr + ch * 64, r);
        tmem_ld_wait();
        if (ch == 1) {                                 // TMEM B is free for dP of the next tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[DQ_FREE]);
        }
#pragma unroll
        for (int c = 0; c < 64; ++c) red_add_f32(dst + (size_t)(ch * 64 + c) * D, __uint_as_float(r[c]) * sm);
      }
    }
  } else if (warp < 12) {
    // =========================== compute: P^T (-> TMEM) and dS^T (-> shared memory) ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 176;");
    const int cw = warp - 4;
    const int quad = cw & 3, half = cw >> 2;
    const int row = quad * 32 + lane;                  // key inside the tile = TMEM lane
    const uint32_t lane_addr = tbase + ((uint32_t)(quad * 32) << 16);
    const uint32_t ds_base = smem_u32(smem) + L::off_ds + half * kAtom;
    const float qk = p.qk_scale;
    for (int n = 0; n < nt; ++n) {
      const int st = n & 1;
      const uint32_t ph = n & 1;
      const bool diag = p.causal && (n == 0);          // i == j: the only tile that needs the mask
      const float* lse_t = ld_s + st * 256 + half * 64;
      const float* dl_t = lse_t + 128;
      float pf[64];
      mbar_wait(&bars[Q_FULL0 + st], (n >> 1) & 1);    // lse / delta of this tile have landed
      mbar_wait(&bars[S_FULL], ph);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t rs[32], w[16];
        tmem_ld32(lane_addr + tA + half * 64 + ch * 32, rs);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 l4 = *reinterpret_cast<const float4*>(lse_t + ch * 32 + c);
          float e0 = ex2_approx(fmaf(__uint_as_float(rs[c + 0]), qk, -l4.x));     // attention_bf16.py:391-392
          float e1 = ex2_approx(fmaf(__uint_as_float(rs[c + 1]), qk, -l4.y));
          float e2 = ex2_approx(fmaf(__uint_as_float(rs[c + 2]), qk, -l4.z));
          float e3 = ex2_approx(fmaf(__uint_as_float(rs[c + 3]), qk, -l4.w));
          if (diag) {                                  // strict causal: key < query keeps its weight (row 0 of the head: fixup kernel)
            const int q0 = half * 64 + ch * 32 + c;
            if (row >= q0 + 0) e0 = 0.f;
            if (row >= q0 + 1) e1 = 0.f;
            if (row >= q0 + 2) e2 = 0.f;
            if (row >= q0 + 3) e3 = 0.f;
          }
          pf[ch * 32 + c + 0] = e0; pf[ch * 32 + c + 1] = e1; pf[ch * 32 + c + 2] = e2; pf[ch * 32 + c + 3] = e3;
          __nv_bfloat162 b0 = __floats2bfloat162_rn(e0, e1), b1 = __floats2bfloat162_rn(e2, e3);
          w[c / 2 + 0] = *reinterpret_cast<uint32_t*>(&b0);
          w[c / 2 + 1] = *reinterpret_cast<uint32_t*>(&b1);
        }
        tmem_st16(lane_addr + tA + half * 64 + ch * 16, w);   // bf16 pairs over the S^T columns this warp has already read
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[P_READY]);

      mbar_wait(&bars[DP_FULL], ph);
      tc_fence_after();
      if (n > 0) mbar_wait(&bars[DS_FREE], (n - 1) & 1);       // dQ / dK of the previous tile have read the dS^T buffer
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t rp[32];
        tmem_ld32(lane_addr + tB + half * 64 + ch * 32, rp);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t wd[4];
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const int c = g * 8 + e * 2;
            const float4 d4 = *reinterpret_cast<const float4*>(dl_t + ch * 32 + c);
            const float s0 = pf[ch * 32 + c + 0] * (__uint_as_float(rp[c + 0]) - d4.x);     // dS = P * (dP - delta)
            const float s1 = pf[ch * 32 + c + 1] * (__uint_as_float(rp[c + 1]) - d4.y);
            const float s2 = pf[ch * 32 + c + 2] * (__uint_as_float(rp[c + 2]) - d4.z);
            const float s3 = pf[ch * 32 + c + 3] * (__uint_as_float(rp[c + 3]) - d4.w);
            __half2 h0 = __floats2half2_rn(s0, s1), h1 = __floats2half2_rn(s2, s3);
            wd[e + 0] = *reinterpret_cast<uint32_t*>(&h0);
            wd[e + 1] = *reinterpret_cast<uint32_t*>(&h1);
          }
          sts128(ds_base + swz128(row, (ch * 32 + g * 8) * 2), wd[0], wd[1], wd[2], wd[3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[DS_READY]);
    }
    // ---- epilogue: dV_j, dK_j accumulators from TMEM (lane = key)
    mbar_wait(&bars[ACC_FULL], 0);
    tc_fence_after();
    const size_t krow = head_row0 + (size_t)j * 128 + row;
    float* dv_dst = p.dv + krow * D + half * 64;
    float* dk_dst = p.dk + krow * D + half * 64;
    const float sm = p.sm_scale;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t r[32];
      tmem_ld32(lane_addr + tDV + half * 64 + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        *reinterpret_cast<float4*>(dv_dst + ch * 32 + c) =
            make_float4(__uint_as_float(r[c]), __uint_as_float(r[c + 1]), __uint_as_float(r[c + 2]), __uint_as_float(r[c + 3]));
      tmem_ld32(lane_addr + tDK + half * 64 + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; c += 4)
        *reinterpret_cast<float4*>(dk_dst + ch * 32 + c) =
            make_float4(__uint_as_float(r[c]) * sm, __uint_as_float(r[c + 1]) * sm, __uint_as_float(r[c + 2]) * sm,
                        __uint_as_float(r[c + 3]) * sm);
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 12) {
      // =========================== tcgen05.mma issue ===========================
      if (elect_one()) {
        constexpr uint32_t id_s = umma_idesc(1, 0, 0, 0, 0, 128, 128);      // S^T : f16, K (K-major) x Q (K-major)
        constexpr uint32_t id_dp = umma_idesc(1, 1, 1, 0, 0, 128, 128);     // dP^T: bf16, V x dO
        constexpr uint32_t id_dv = umma_idesc(1, 1, 1, 0, 1, 128, D);       // dV  : bf16, P^T (TMEM) x dO (MN-major)
        constexpr uint32_t id_dk = umma_idesc(1, 0, 0, 0, 1, 128, D);       // dK  : f16, dS^T (K-major) x Q (MN-major)
        constexpr uint32_t id_dq = umma_idesc(1, 0, 0, 1, 1, 128, 128);     // dQ^T: f16, K^T (MN-major) x dS^T (MN-major)
        const uint32_t a_k = smem_u32(smem + L::off_k), a_v = smem_u32(smem + L::off_v), a_ds = smem_u32(smem + L::off_ds);
        auto issue_s = [&](int st) {
          const uint32_t a_q = smem_u32(smem + L::off_q + st * L::kTile);
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t o = (k >> 2) * kAtom + (k & 3) * 32;
            umma_f16_ss(tbase + tA, umma_smem_desc(a_k + o, 16, 1024, kSwz128), umma_smem_desc(a_q + o, 16, 1024, kSwz128), id_s, k > 0);
          }
          umma_commit(&bars[S_FULL]);
        };
        auto issue_dp = [&](int st) {
          const uint32_t a_do = smem_u32(smem + L::off_do + st * L::kTile);
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t o = (k >> 2) * kAtom + (k & 3) * 32;
            umma_f16_ss(tbase + tB, umma_smem_desc(a_v + o, 16, 1024, kSwz128), umma_smem_desc(a_do + o, 16, 1024, kSwz128), id_dp, k > 0);
          }
          umma_commit(&bars[DP_FULL]);
        };
        mbar_wait(&bars[KV_FULL], 0);
        mbar_wait(&bars[Q_FULL0], 0);
        tc_fence_after();
        issue_s(0);
        mbar_wait(&bars[DO_FULL0], 0);
        issue_dp(0);
        for (int n = 0; n < nt; ++n) {
          const int st = n & 1;
          const uint32_t a_q = smem_u32(smem + L::off_q + st * L::kTile), a_do = smem_u32(smem + L::off_do + st * L::kTile);
          mbar_wait(&bars[P_READY], n & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k)                           // contraction over the 128 queries, 16 per instruction
            umma_f16_ts(tbase + tDV, tbase + tA + (k >> 2) * 64 + (k & 3) * 8, umma_smem_desc(a_do + k * 2048, kAtom, 1024, kSwz128),
                        id_dv, (n > 0) || (k > 0));
          umma_commit(&bars[DO_FREE0 + st]);
          if (n + 1 < nt) {
            mbar_wait(&bars[Q_FULL0 + (st ^ 1)], ((n + 1) >> 1) & 1);
            tc_fence_after();
            issue_s(st ^ 1);                                    // overwrites P^T(n): tcgen05.mma of one thread execute in order
          }
          mbar_wait(&bars[DS_READY], n & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k)                           // contraction over the 128 keys
            umma_f16_ss(tbase + tB, umma_smem_desc(a_k + k * 2048, kAtom, 1024, kSwz128),
                        umma_smem_desc(a_ds + k * 2048, kAtom, 1024, kSwz128), id_dq, k > 0);
          umma_commit(&bars[DQ_FULL]);
#pragma unroll
          for (int k = 0; k < 8; ++k)                           // contraction over the 128 queries
            umma_f16_ss(tbase + tDK, umma_smem_desc(a_ds + (k >> 2) * kAtom + (k & 3) * 32, 16, 1024, kSwz128),
                        umma_smem_desc(a_q + k * 2048, kAtom, 1024, kSwz128), id_dk, (n > 0) || (k > 0));
          umma_commit(&bars[Q_FREE0 + st]);
          umma_commit(&bars[DS_FREE]);
          if (n + 1 < nt) {
            mbar_wait(&bars[DO_FULL0 + (st ^ 1)], ((n + 1) >> 1) & 1);
            mbar_wait(&bars[DQ_FREE], n & 1);
            tc_fence_after();
            issue_dp(st ^ 1);
          }
        }
        umma_commit(&bars[ACC_FULL]);
      }
      __syncwarp();
    } else if (warp == 13) {
      // =========================== TMA issue ===========================
      if (elect_one()) {
        mbar_expect_tx(&bars[KV_FULL], 2 * L::kTile);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          tma_load_2d(smem + L::off_k + a * kAtom, &tm_k, &bars[KV_FULL], a * 64, (int)head_row0 + j * 128);
          tma_load_2d(smem + L::off_v + a * kAtom, &tm_v, &bars[KV_FULL], a * 64, (int)head_row0 + j * 128);
        }
        for (int n = 0; n < nt; ++n) {
          const int st = n & 1;
          const int r0 = (int)head_row0 + (i0 + n) * 128;
          if (n >= 2) mbar_wait(&bars[Q_FREE0 + st], ((n >> 1) - 1) & 1);
          mbar_expect_tx(&bars[Q_FULL0 + st], L::kTile + 1024);
#pragma unroll
          for (int a = 0; a < 2; ++a) tma_load_2d(smem + L::off_q + st * L::kTile + a * kAtom, &tm_q, &bars[Q_FULL0 + st], a * 64, r0);
          bulk_load_1d(ld_s + st * 256, p.lse + r0, 512, &bars[Q_FULL0 + st]);
          bulk_load_1d(ld_s + st * 256 + 128, p.delta + r0, 512, &bars[Q_FULL0 + st]);
          if (n >= 2) mbar_wait(&bars[DO_FREE0 + st], ((n >> 1) - 1) & 1);
          mbar_expect_tx(&bars[DO_FULL0 + st], L::kTile);
#pragma unroll
          for (int a = 0; a < 2; ++a) tma_load_2d(smem + L::off_do + st * L::kTile + a * kAtom, &tm_do, &bars[DO_FULL0 + st], a * 64, r0);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc<512>(tbase);
}

int launch_bf16_bwd_ws(const void* q, const void* k, const void* v, const void* do_bf16, const float* lse, const float* delta,
                       float* dq, float* dk, float* dv, int BH, int S, int causal, cudaStream_t st) {
  using L = Bf16Bwd2Smem;
  constexpr int D = 128;
  CUtensorMap tq, tk, tv, tdo;
  uint64_t dims[2] = {(uint64_t)D, (uint64_t)BH * S};
  uint64_t str[1] = {(uint64_t)D * 2};
  uint32_t box[2] = {64, 128};
  int rc;
  if ((rc = qa_make_tmap(&tq, q, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tk, k, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tv, v, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tdo, do_bf16, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  Bf16BwdParams2 p;
  p.lse = lse; p.delta = delta; p.dq = dq; p.dk = dk; p.dv = dv; p.S = S; p.causal = causal;
  p.sm_scale = (float)(1.0 / sqrt((double)D));
  p.qk_scale = (float)((1.0 / sqrt((double)D)) * 1.44269504);
  cudaError_t e = cudaFuncSetAttribute(bf16_bwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  dim3 grid(S / 128, BH);
  bf16_bwd_ws_kernel<<<grid, 512, L::total, st>>>(tq, tk, tv, tdo, p);
  return qa_check_launch("qa_bf16_bwd(ws)");
}

}  // namespace qa
