// bf16 bias-corrected flash attention forward, two-query-tile variant (used when Sq % 256 == 0).
// Same numerics as attn_bf16_fwd.cu (reference attention_bf16.py:195-294, contract mode); different schedule:
//   * one CTA owns TWO 128-row query tiles (A, B) of one head and shares every K/V tile between them;
//   * each softmax warpgroup owns one query tile, one thread per row (128 columns): no cross-warp exchange;
//   * the softmax advances in steps of 64 keys (half a K/V tile).  S_A / S_B are double-buffered in TMEM
//     (2 x 64 columns each): Q K^T of step t+2 is issued right behind P V of step t, so the logits of step t+1 are
//     already waiting when a softmax warp finishes step t - the softmax warps never idle on the tensor pipe;
//   * P is written back to TMEM over the S columns (bf16, 2 per column) and consumed as the A operand of the
//     P V MMA straight from TMEM (no shared-memory round trip, no proxy fence);
//   * O_A / O_B stay resident in TMEM; a softmax warp rescales its own 32 rows only when one of their maxima moved.
//     With the lazy running maximum (rescale_tau, qa_bf16_fwd_ex) that is rare, so there is no correction warpgroup:
//     10 warps per CTA leave 168 registers per thread for instruction-level parallelism in the softmax loop.
// TMEM (512 columns): S_A[2] [0,64) [64,128)  S_B[2] [128,192) [192,256)  O_A [256,256+D)  O_B [384,384+D).
#include "qa_ptx.cuh"
#include "qa_host.h"
#include <type_traits>
#include <stdlib.h>

namespace qa {

constexpr int kAtom2 = 128 * 128;

template <int D, int STAGES>
struct Bf16Fwd2Smem {
  static constexpr int kTile = 128 * D * 2;
  static constexpr int off_q = 0;                          // Q_A, Q_B
  static constexpr int off_k = off_q + 2 * kTile;
  static constexpr int off_v = off_k + STAGES * kTile;
  static constexpr int total = off_v + STAGES * kTile + 1024;
};

struct Bf16Fwd2Params {
  float* O;
  float* lse;
  int Sq, Sk, causal, BH;
  int Sk_valid;        // keys [Sk_valid, Sk) of every head are zero padding (ragged sequence): weight exactly 0
  float qk_scale;
  float rescale_tau;   // adopt a new running maximum only when it exceeds the current one by more than this (log2 units)
  long long* dbg;      // development library only: [CTA][16] globaltimer stamps (tools/timeline_bf16_fwd.py)
};

#ifdef QA_DEV_TIMELINE
#define QA_TLF(slot)                                                                                      \
  do {                                                                                                    \
    if (p.dbg != nullptr) {                                                                               \
      long long t_;                                                                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                              \
      p.dbg[(size_t)blockIdx.x * 16 + (slot)] = t_;                            \
    }                                                                                                     \
  } while (0)
#else
#define QA_TLF(slot) do { } while (0)
#endif

__device__ __forceinline__ float bf2_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf2_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ __nv_bfloat162 u2bf(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }
__device__ __forceinline__ uint32_t bf2u(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

template <int D, int STAGES, int KS>
__global__ void __launch_bounds__(576, 1)
bf16_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, Bf16Fwd2Params p) {
  using L = Bf16Fwd2Smem<D, STAGES>;
  constexpr int kDAtoms = D / 64;
  constexpr int NB = 128 / KS;                                   // S / P buffers per query tile: steps of KS keys
  static_assert(KS == 64 || KS == 128, "keys per softmax step");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t q_full, k_full[STAGES], k_empty[STAGES], v_full[STAGES], v_empty[STAGES];
  __shared__ uint64_t s_full[2][2], p_full[2][2], o_full[2][2], o_ready[2][2];   // [query tile][S buffer]: a softmax warp may run up to two steps ahead
                                                                                   // of the issuer, so every barrier is per buffer (a single one aliases parities)
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t xtop[2][2][2][128];      // [query tile][step parity][column half][row]: packed (top1, top2) bf16
  __shared__ float xl[2][128];                 // row sums of column half 1

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npair = p.Sq / 256;
  int rank, bh;
  qa_group_order((int)blockIdx.x, p.BH, npair, p.causal ? 16 : 1, rank, bh);
  const int pt = npair - 1 - rank;                                // pair of query tiles; heaviest causal pairs first
  const int q0 = pt * 256;
  const int nk_full = (p.Sk_valid + 127) / 128;                   // k-tiles without a valid key are skipped
  // per query tile: number of k-tiles that contain at least one visible key (strict causal: key < query)
  const int nkx[2] = {p.causal ? min(nk_full, 2 * pt + 1) : nk_full, p.causal ? min(nk_full, 2 * pt + 2) : nk_full};
  const int nk = nkx[1];

  if (tid == 0) {
    QA_TLF(0);
#ifdef QA_DEV_TIMELINE
    if (p.dbg != nullptr) { uint32_t sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); p.dbg[(size_t)blockIdx.x * 16 + 15] = sm_; }
#endif
    mbar_init(&q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&s_full[x][0], 1); mbar_init(&s_full[x][1], 1); mbar_init(&p_full[x][0], 8); mbar_init(&p_full[x][1], 8);
      mbar_init(&o_full[x][0], 1); mbar_init(&o_full[x][1], 1); mbar_init(&o_ready[x][0], 8); mbar_init(&o_ready[x][1], 8);
    }
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  if (tid == 0) QA_TLF(1);

  if (warp < 16) {
    // =========================== softmax warps: 0-7 -> tile A, 8-15 -> tile B; two warps per 32-row group, ===========================
    // each with half of the step's key columns (one warp per row was the limiter once S moved to N = 128 instructions: the
    // softmax of one tile has to fit in the tensor-pipe time of the other tile)
    const int x = warp >> 3, hf = (warp >> 2) & 1, qd = warp & 3;
    constexpr int KH = KS / 2;                                     // key columns per warp and step
    const int row = qd * 32 + lane;
    const uint32_t s_addr = tbase + ((uint32_t)(qd * 32) << 16) + x * 128;
    const uint32_t o_addr = tbase + ((uint32_t)(qd * 32) << 16) + 256 + x * 128;
    const int grow = q0 + x * 128 + row;                          // query index inside the head
    const int qt = 2 * pt + x;                                    // this tile's diagonal k-tile
    const int nst = NB * nkx[x];                                  // KS-key steps of this query tile
    __nv_bfloat16 m_bf = __float2bfloat16(-INFINITY);
    float l = hf == 0 ? 1.0f : 0.0f;                              // attention_bf16.py:198 (the two column halves add up)
    const uint32_t ninf2 = 0xff80ff80u;
    const float2 qk2 = make_float2(p.qk_scale, p.qk_scale);
    for (int t = 0; t < nst; ++t) {
      const int j = t / NB, b = t % NB;
      const uint32_t bph = (t / NB) & 1;                           // phase of this step's barriers (one per buffer)
      const bool diag = p.causal && (j == qt);
      const bool tail = (t + 1) * KS > p.Sk_valid;                 // ragged: this step holds padding
      const int klim = diag ? min(grow, p.Sk_valid) : p.Sk_valid;  // first key without weight
      const uint32_t sb_addr = s_addr + b * KS;
      mbar_wait(&s_full[x][b], bph);
      tc_fence_after();
      if (lane == 0 && qd == 0 && hf == 0 && t == 0) QA_TLF(2 + x);
      if (lane == 0 && qd == 0 && hf == 0 && t == 4) QA_TLF(4 + x);
      // ---- pass 1: u = bf16(S * qk_scale), masked; top-2 of this warp's column half
      uint32_t u2[KH / 2];
      __nv_bfloat162 t1 = u2bf(ninf2), t2 = u2bf(ninf2);
      auto pass1 = [&](auto masked) {                              // two instantiations: the mask costs nothing off-diagonal
#pragma unroll
        for (int ch = 0; ch < KH / 32; ++ch) {
          uint32_t r[32];
          tmem_ld32(sb_addr + hf * KH + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const __nv_bfloat162 ub = __float22bfloat162_rn(__fmul2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), qk2));
            uint32_t u = bf2u(ub);                                 // u = bf16(S * qk_scale), one FMUL2 + one pack per pair
            if (decltype(masked)::value) {                         // strict causal: keep key < query; ragged: key < Sk_valid
              const int key = t * KS + hf * KH + ch * 32 + 2 * i;
              if (key >= klim) u = (u & 0xffff0000u) | 0xff80u;
              if (key + 1 >= klim) u = (u & 0x0000ffffu) | 0xff800000u;
            }
            u2[ch * 16 + i] = u;
            const __nv_bfloat162 xv = u2bf(u);
            t2 = __hmax2(t2, __hmin2(t1, xv));
            t1 = __hmax2(t1, xv);
          }
        }
      };
      if (diag || tail) pass1(std::true_type{}); else pass1(std::false_type{});
      const __nv_bfloat16 a1 = __low2bfloat16(t1), b1 = __high2bfloat16(t1), a2 = __low2bfloat16(t2), b2 = __high2bfloat16(t2);
      __nv_bfloat16 top1 = __hmax(a1, b1);
      __nv_bfloat16 top2 = __hmax(__hmin(a1, b1), __hmax(a2, b2));
      {                                                            // top-2 of the whole row: exchange with the warp of the other half
        __nv_bfloat162 mine = __halves2bfloat162(top1, top2);
        xtop[x][t & 1][hf][row] = bf2u(mine);
        named_bar_sync(1 + x * 4 + qd, 64);
        const __nv_bfloat162 oth = u2bf(xtop[x][t & 1][hf ^ 1][row]);
        const __nv_bfloat16 o1 = __low2bfloat16(oth), o2 = __high2bfloat16(oth);
        top2 = __hmax(__hmin(top1, o1), __hmax(top2, o2));
        top1 = __hmax(top1, o1);
      }
      // ---- bias-corrected running max (attention_bf16.py:236-264, predicate in the scaled domain)
      __nv_bfloat16 m_new = __hmax(m_bf, top1);
      const __nv_bfloat16 thr = __float2bfloat16(__bfloat162float(m_new) - 1e-3f);
      const bool many = (top2 >= thr) && (top1 >= thr);
      const float mf = __bfloat162float(m_new);
      if (many && mf > 0.f) m_new = __float2bfloat16(2.0f * mf);
      else if (many && mf < 0.f) m_new = __float2bfloat16(0.f);
      // lazy rescale: while the candidate is within rescale_tau of the current maximum keep the current one
      // (P <= 2^tau, mathematically neutral; rescale_tau = 0 is the step-by-step maximum of the reference)
      if (!(__bfloat162float(__hsub(m_new, m_bf)) > p.rescale_tau)) m_new = m_bf;
      const float resc = __bfloat162float(__float2bfloat16(ex2_approx(__bfloat162float(__hsub(m_bf, m_new)))));
      m_bf = m_new;
      if (t > 0 && __any_sync(0xffffffffu, resc != 1.0f)) {       // O *= rescale (:280): rare with the lazy maximum
        mbar_wait(&o_full[x][(t - 1) % NB], ((t - 1) / NB) & 1); // P V of step t-1 has landed in TMEM
        tc_fence_after();
        const float2 rs2 = make_float2(resc, resc);
#pragma unroll
        for (int ch = 0; ch < D / 64; ++ch) {                      // this warp's half of the D columns
          uint32_t r[32];
          tmem_ld32(o_addr + hf * (D / 2) + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 o2 = __fmul2_rn(make_float2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), rs2);
            r[i] = __float_as_uint(o2.x); r[i + 1] = __float_as_uint(o2.y);
          }
          tmem_st32(o_addr + hf * (D / 2) + ch * 32, r);
        }
        tmem_st_wait();
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_ready[x][b]);                 // the MMA warp may accumulate step t into O
      // ---- pass 2: P = bf16(exp2(bf16(u - m'))) written back over the S columns (2 per column), l += sum(P).  Both warps
      //      of the row group are past pass 1 (the exchange above), so the P columns of the other half may be overwritten.
      const __nv_bfloat162 m2 = __bfloat162bfloat162(m_new);
      float2 ls2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int g = 0; g < KH / 64 + (KH < 64 ? 1 : 0); ++g) {
        constexpr int NW = KH < 64 ? KH / 2 : 32;
        uint32_t w[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const uint32_t d = bf2u(__hsub2(u2bf(u2[g * 32 + i]), m2));
          const uint32_t pp = pack2_bf16(ex2_approx(bf2_lo(d)), ex2_approx(bf2_hi(d)));
          ls2 = __fadd2_rn(ls2, make_float2(bf2_lo(pp), bf2_hi(pp)));
          w[i] = pp;
        }
        if constexpr (NW == 32) tmem_st32(sb_addr + hf * (KH / 2) + g * 32, w);
        else tmem_st16(sb_addr + hf * (KH / 2), w);
      }
      tmem_st_wait();
      l = l * resc + (ls2.x + ls2.y);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[x][b]);
    }
    if (lane == 0 && qd == 0 && hf == 0) QA_TLF(6 + x);
    // ---- epilogue (the warps of column half 0): O / l and the log2-LSE, straight from the resident accumulator
    if (hf == 1) xl[x][row] = l;
    named_bar_sync(1 + x * 4 + qd, 64);
    if (hf == 0) {
    l += xl[x][row];
    mbar_wait(&o_full[x][(nst - 1) % NB], ((nst - 1) / NB) & 1);
    tc_fence_after();
    if (lane == 0 && qd == 0) QA_TLF(8 + x);
    const size_t gr = (size_t)bh * p.Sq + grow;
    const float inv_l = 1.0f / l;
    // O rows go out through TMA stores: each warp stages its 32 rows as [32 rows][32 floats] swizzled slices (4 KB) in the
    // shared memory of this tile's Q (every S MMA of the tile has completed), two rounds
    constexpr int SL = L::kTile / 16384;                          // slices per warp and round
    const int w4 = qd;
#pragma unroll
    for (int rd = 0; rd < (D / 32) / SL; ++rd) {
      if (rd > 0) {
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
#pragma unroll
      for (int sl = 0; sl < SL; ++sl) {
        uint32_t r[32];
        tmem_ld32(o_addr + (rd * SL + sl) * 32, r);
        tmem_ld_wait();
        const uint32_t base = smem_u32(smem) + L::off_q + x * L::kTile + (sl * 4 + w4) * 4096;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          sts128f(base + swz128(lane, i * 4), __uint_as_float(r[i]) * inv_l, __uint_as_float(r[i + 1]) * inv_l,
                  __uint_as_float(r[i + 2]) * inv_l, __uint_as_float(r[i + 3]) * inv_l);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int sl = 0; sl < SL; ++sl)
          tma_store_2d(&tm_o, smem + L::off_q + x * L::kTile + (sl * 4 + w4) * 4096, (rd * SL + sl) * 32,
                       bh * p.Sq + q0 + x * 128 + w4 * 32);
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_read();
    p.lse[gr] = __bfloat162float(m_bf) + log2f(l);                             // attention_bf16.py:288
    if (lane == 0 && qd == 0) QA_TLF(10 + x);
    }
  } else if (warp == 16) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
      mbar_expect_tx(&q_full, 2 * L::kTile);
#pragma unroll
      for (int x = 0; x < 2; ++x)
#pragma unroll
        for (int a = 0; a < kDAtoms; ++a)
          tma_load_2d(smem + L::off_q + x * L::kTile + a * kAtom2, &tm_q, &q_full, a * 64, bh * p.Sq + q0 + x * 128);
      for (int j = 0; j < nk; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], L::kTile);
#pragma unroll
        for (int a = 0; a < kDAtoms; ++a)
          tma_load_2d(smem + L::off_k + s * L::kTile + a * kAtom2, &tm_k, &k_full[s], a * 64, bh * p.Sk + j * 128);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], L::kTile);
#pragma unroll
        for (int a = 0; a < kDAtoms; ++a)
          tma_load_2d(smem + L::off_v + s * L::kTile + a * kAtom2, &tm_v, &v_full[s], a * 64, bh * p.Sk + j * 128);
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc(1, 0, 0, 0, 0, 128, KS);         // f32 += f16 x f16, K-major, KS keys
      constexpr uint32_t idesc_pv = umma_idesc(1, 1, 1, 0, 1, 128, D);          // f32 += bf16 (TMEM) x bf16 (V MN-major)
      auto issue_qk = [&](int x, int t) {                                        // S_x[t % NB] = Q_x K_step(t)^T
        const int s = (t / NB) % STAGES, h = t % NB;
        const uint32_t q_addr = smem_u32(smem + L::off_q + x * L::kTile);
        const uint32_t k_addr = smem_u32(smem + L::off_k + s * L::kTile) + h * (KS * 128);   // rows h * KS .. of every 64-column atom
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t o = (k >> 2) * kAtom2 + (k & 3) * 32;
          umma_f16_ss(tbase + x * 128 + h * KS, umma_smem_desc(q_addr + o, 16, 1024, kSwz128), umma_smem_desc(k_addr + o, 16, 1024, kSwz128),
                      idesc_qk, k > 0);
        }
        umma_commit(&s_full[x][h]);
      };
      auto issue_pv = [&](int x, int t) {                                        // O_x += P_x V_step(t), P from TMEM
        const int s = (t / NB) % STAGES, h = t % NB;
        mbar_wait(&p_full[x][h], (t / NB) & 1);
        mbar_wait(&o_ready[x][h], (t / NB) & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(smem + L::off_v + s * L::kTile) + h * ((KS / 16) * 2048);
#pragma unroll
        for (int k = 0; k < KS / 16; ++k)
          umma_f16_ts(tbase + 256 + x * 128, tbase + x * 128 + h * KS + k * 8, umma_smem_desc(v_addr + k * 2048, kAtom2, 1024, kSwz128),
                      idesc_pv, (t > 0) || (k > 0));
        umma_commit(&o_full[x][h]);
      };
      const int nst[2] = {NB * nkx[0], NB * nkx[1]};
      mbar_wait(&q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        issue_qk(0, b);
        issue_qk(1, b);
      }
      umma_commit(&k_empty[0]);
      for (int t = 0; t < NB * nk; ++t) {
        const int j = t / NB, h = t % NB;
        const bool more = (j + 1 < nk);
        if (h == 0) {
          mbar_wait(&v_full[j % STAGES], (j / STAGES) & 1);
          if (more) mbar_wait(&k_full[(j + 1) % STAGES], ((j + 1) / STAGES) & 1);
          tc_fence_after();
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          if (t < nst[x]) {
            issue_pv(x, t);
            if (t + NB < nst[x]) issue_qk(x, t + NB);             // behind P V in the in-order pipe: P_x(t) is consumed first
          }
        }
        if (h == NB - 1) {
          umma_commit(&v_empty[j % STAGES]);
          if (more) umma_commit(&k_empty[(j + 1) % STAGES]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<512>(tbase);
  if (tid == 0) QA_TLF(12);
}

#ifdef QA_DEV_TIMELINE
static void* g_bf16_fwd_dbg = nullptr;
// Development library only (include/qattn_dev.h, tools/timeline_bf16_fwd.py): per-CTA globaltimer stamps of subsequent
// qa_bf16_fwd launches of the two-tile kernel ([CTAs][16] int64); NULL = off.
extern "C" int qa_debug_set_bf16_fwd_timeline(void* buf) {
  g_bf16_fwd_dbg = buf;
  return 0;
}
#endif

template <int D, int STAGES>
int launch_bf16_fwd2(const void* q, const void* k, const void* v, float* O, float* lse, int BH, int Sq, int Sk, int Sk_valid,
                     int causal, float qk_scale, float rescale_tau, cudaStream_t st) {

  using L = Bf16Fwd2Smem<D, STAGES>;
  CUtensorMap tq, tk, tv, to;
  uint64_t dq[2] = {(uint64_t)D, (uint64_t)BH * Sq}, dk[2] = {(uint64_t)D, (uint64_t)BH * Sk};
  uint64_t str[1] = {(uint64_t)D * 2};
  uint32_t box[2] = {64, 128};
  uint64_t stro[1] = {(uint64_t)D * 4};
  uint32_t boxo[2] = {32, 32};                         // one warp's 32 rows x 32 fp32 columns of O
  int rc;
  if ((rc = qa_make_tmap(&to, O, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dq, stro, boxo, 3))) return rc;
  if ((rc = qa_make_tmap(&tq, q, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dq, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tk, k, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dk, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tv, v, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dk, str, box, 3))) return rc;
  Bf16Fwd2Params p;
  p.O = O; p.lse = lse; p.Sq = Sq; p.Sk = Sk; p.Sk_valid = Sk_valid; p.causal = causal; p.BH = BH; p.qk_scale = qk_scale; p.rescale_tau = rescale_tau;
#ifdef QA_DEV_TIMELINE
  p.dbg = (long long*)g_bf16_fwd_dbg;
#else
  p.dbg = nullptr;
#endif
  auto kern = bf16_fwd2_kernel<D, STAGES, 128>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  dim3 grid((Sq / 256) * BH);                         // order: qa_group_order
  kern<<<grid, 576, L::total, st>>>(tq, tk, tv, to, p);
  return qa_check_launch("qa_bf16_fwd(2 query tiles)");
}

template int launch_bf16_fwd2<128, 2>(const void*, const void*, const void*, float*, float*, int, int, int, int, int, float, float, cudaStream_t);
template int launch_bf16_fwd2<64, 3>(const void*, const void*, const void*, float*, float*, int, int, int, int, int, float, float, cudaStream_t);

}  // namespace qa
