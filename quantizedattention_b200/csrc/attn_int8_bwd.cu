// int8 SageAttention3-style backward (reference attention_int8.py:342-428 with the
// 8-LEDGER contract: dS = P*(dP - delta), sm_scale, accumulation over tiles, fp32 delta pre-pass, fp32 lse).
//
// One CTA = one 128-key tile of one (batch, head), looping over the 128-row query tiles (k-outer, as the
// reference).  Five tcgen05 kind::i8 contractions per tile pair:
//   S    = Q_i  K_j^T     dP    = dO_i V_j^T         (phase A, TMEM cols 0..127 / 128..255)
//   dV_j += P^T dO_i      dK_j += dS^T Q_i     dQ_i += dS K_j   (phase C, TMEM 256.. / 384.. / 0.. aliasing S)
// P and dS are re-quantised per [128 x 128] tile (scale = amax/127, truncation), so every int32 product is
// drained to fp32 after each tile pair: dV/dK accumulate in registers, the dQ tile is staged in shared memory and
// added to an fp32 workspace by a TMA reduce-add (cp.reduce.async.bulk.tensor); the summation order over k-tiles is
// therefore not deterministic (LEDGER I-9).  P / dS are computed twice (amax pass, quantise pass) from the packed
// fp16 logits instead of being held in registers, which keeps the kernel spill-free.
// Phases are separated by named barriers + mbarriers; 8 warps compute (thread = row, half of the columns) and
// thread 0 issues TMA and MMA at the phase boundaries.
#include "qa_ptx.cuh"
#include "qa_host.h"
#include <type_traits>

namespace qa {

template <int D>
struct Int8BwdSmem {
  static constexpr int kTile = 128 * D;              // int8 [128][D]
  static constexpr int off_k = 0;
  static constexpr int off_v = off_k + kTile;
  static constexpr int off_q = off_v + kTile;        // 2 stages
  static constexpr int off_do = off_q + 2 * kTile;   // 2 stages
  static constexpr int off_p = off_do + 2 * kTile;   // int8 [128 q][128 keys]
  static constexpr int off_ds = off_p + 128 * 128;   // 2 buffers (dQ of tile t-1 reads dS while tile t is quantised)
  static constexpr int off_dq = off_ds + 2 * 128 * 128;  // fp32 [128 q][D] staging for the TMA reduce-add (D/32 swizzled atoms)
  static constexpr int off_c = off_dq + 128 * D * 4;     // two 1 KB constant fp16 atoms (A: 1024.0, B: 768.0) of the accumulator-initialising MMA
  static constexpr int total = off_c + 2 * 1024 + 1024;
};

// Shared memory of the block kernel in SageBwd mode (dP = dO V^T kept in fp16): no int8 V tile, one stage
// of Q / dO, one dS buffer, plus the fp16 V tile of this k-tile and the fp16 dO tile of the current query tile
// ([128 rows][64 columns] 128B-swizzled atoms, as the bf16 kernels lay out their operands).
template <int D>
struct Int8BwdSageSmem {
  static constexpr int kTile = 128 * D;              // int8 [128][D]
  static constexpr int kTile16 = 128 * D * 2;        // fp16 [128][D] as D/64 atoms of 16 KB
  static constexpr int off_k = 0;
  static constexpr int off_q = off_k + kTile;
  static constexpr int off_do = off_q + kTile;
  static constexpr int off_p = off_do + kTile;
  static constexpr int off_ds = off_p + 128 * 128;
  static constexpr int off_dq = off_ds + 128 * 128;  // fp32 [128 q][D] staging for the TMA reduce-add
  static constexpr int off_c = off_dq + 128 * D * 4;
  static constexpr int off_v16 = off_c + 2 * 1024;
  static constexpr int off_do16 = off_v16 + kTile16;
  static constexpr int total = off_do16 + kTile16 + 1024;
  static constexpr int off_v = off_k;                // unused in this mode (kept so that shared code compiles)
};

struct Int8BwdParams {
  const __half *sq, *sk, *sv, *s_do;   // per-128-block fp16 scales
  const float* lse;                    // [BH*S] fp32 log2-sum-exp2
  const float* delta;                  // [BH*S]
  float* rowsum;                       // [BH*S] fp32, zero-initialised: sum over k-tiles of rowsum(dS) (null without K smoothing)
  __half *dk, *dv;                     // [BH*S, D] fp16
  int S;
  int S_valid;                         // rows [S_valid, S) of every head are padding (ragged sequence): padded keys get P = 0
  float sm_scale, qk_scale;
  long long* dbg;                      // optional timeline buffer [tile][2 warps][16] of SM clock stamps (tools/timeline_bwd.py)
};

// 256 threads: thread = (row, column half).  Thread 0 additionally issues TMA / tcgen05.mma at the two barriers per
// tile.  Software pipeline (tile t): the S/dP MMAs of tile t+1 and the dV/dK MMAs of tile t are issued together, so
// pass 1 of tile t+1 runs while the tensor core computes dV_t/dK_t; dQ_t is issued one barrier later into the TMEM
// columns freed by draining dV_t and overlaps the quantise pass of tile t+1.
//   TMEM: [0,128) S   [128,256) dP   [256,384) dV partial, then dQ partial   [384,512) dK partial
#ifdef QA_DEV_TIMELINE
#define QA_TLB(slot)                                                                                          \
  do {                                                                                                        \
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 0 || warp == 5) && t < 64) \
      p.dbg[(t * 2 + (warp != 0)) * 16 + (slot)] = clock64();                                                 \
  } while (0)
#define QA_TLW(role, slot)                                                                                    \
  do {                                                                                                        \
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && t < 64)                        \
      p.dbg[(t * 2 + (role)) * 16 + (slot)] = clock64();                                                      \
  } while (0)
#else
#define QA_TLB(slot) do { } while (0)
#define QA_TLW(role, slot) do { } while (0)
#endif

template <int D, int NG, bool RN, bool CAUSAL>
__global__ void __launch_bounds__(128 * NG, 1)
int8_bwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                const __grid_constant__ CUtensorMap tm_dq, Int8BwdParams p) {
  using L = Int8BwdSmem<D>;
  constexpr int DH = D / NG;                                    // output columns per thread
  constexpr int CW = 128 / NG;                                  // S / dP columns per thread
  constexpr int NT = 128 * NG;                                  // threads
  constexpr int NW = 4 * NG;                                    // warps
  constexpr uint32_t kLay = (D == 128) ? kSwz128 : kSwz64;      // operand rows of D bytes
  constexpr uint32_t kSbo = (D == 128) ? 1024 : 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t kv_full, qdo_full[2], sd_full, parts_full, dq_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red_p[2][NW], red_ds[2][NW];
  __shared__ float rowsum_ds[2][NG][128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool leader = (tid == 0);
  const int bh = blockIdx.y, j = blockIdx.x;
  const int nq = (p.S_valid + 127) / 128;                       // fully padded query tiles are skipped
  const int ktail = p.S_valid - j * 128;                        // valid keys of this k-tile (>= 128 unless it is the ragged last one)
  // CAUSAL (strict mask, key < query): k-tile j meets the query tiles j .. nq-1; `t` below is the position in
  // that sequence (pipeline stages and barrier parities), t0 + t the query tile
  const int t0 = CAUSAL ? j : 0;
  const int nt = nq - t0;
  const size_t head_row0 = (size_t)bh * p.S;

  if (leader) {
    mbar_init(&kv_full, 1); mbar_init(&qdo_full[0], 1); mbar_init(&qdo_full[1], 1);
    mbar_init(&sd_full, 1); mbar_init(&parts_full, 1); mbar_init(&dq_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  if (tid >= 64 && tid < 192) {                                 // constant atoms of the accumulator-initialising MMA
    const uint32_t v2 = tid < 128 ? kMagicElemA2 : kMagicElemB2;
    sts128(smem_u32(smem + L::off_c) + (tid - 64) * 16, v2, v2, v2, v2);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  constexpr uint32_t id_s = umma_idesc(2, 1, 1, 0, 0, 128, 128);     // S, dP: A, B K-major, N = 128
  constexpr uint32_t id_t = umma_idesc(2, 1, 1, 1, 1, 128, D);       // dV, dK: A^T (MN-major), B MN-major, N = D
  constexpr uint32_t id_q = umma_idesc(2, 1, 1, 0, 1, 128, D);       // dQ: A K-major, B MN-major
  // "magic" accumulators (qa_ptx.cuh): a kind::f16 MMA over two constant atoms sets an accumulator to kMagic before the
  // kind::i8 MMAs add to it; all sixteen 8-row groups read the same 1 KB atom (SBO = 0)
  constexpr uint32_t id_c128 = umma_idesc(1, 0, 0, 0, 0, 128, 128), id_cD = umma_idesc(1, 0, 0, 0, 0, 128, D);
  const uint64_t cdesc_a = umma_smem_desc(smem_u32(smem + L::off_c), 16, 0, kLay), cdesc_b = umma_smem_desc(smem_u32(smem + L::off_c + 1024), 16, 0, kLay);
  const uint32_t a_k = smem_u32(smem + L::off_k), a_v = smem_u32(smem + L::off_v);
  const uint32_t a_p = smem_u32(smem + L::off_p), a_ds0 = smem_u32(smem + L::off_ds);

  auto load_qdo = [&](int tile, int st) {
    mbar_expect_tx(&qdo_full[st], 2 * L::kTile);
    tma_load_2d(smem + L::off_q + st * L::kTile, &tm_q, &qdo_full[st], 0, (int)head_row0 + tile * 128);
    tma_load_2d(smem + L::off_do + st * L::kTile, &tm_do, &qdo_full[st], 0, (int)head_row0 + tile * 128);
  };
  auto issue_s = [&](int st) {                                     // S = Q K^T (issued one barrier ahead of dP: the S columns
    const uint32_t a_q = smem_u32(smem + L::off_q + st * L::kTile);          // are free as soon as pass 1 is over)
      umma_f16_ss(tbase + 0, cdesc_a, cdesc_b, id_c128, 0);                     // S = kMagic (qa_ptx.cuh)
#pragma unroll
    for (int k = 0; k < D / 32; ++k)
      umma_i8_ss(tbase + 0, umma_smem_desc(a_q + k * 32, 16, kSbo, kLay), umma_smem_desc(a_k + k * 32, 16, kSbo, kLay), id_s, 1);
  };
  auto issue_dp = [&](int st) {                                    // dP = dO V^T; the commit also covers the earlier S MMAs
    const uint32_t a_do = smem_u32(smem + L::off_do + st * L::kTile);
      umma_f16_ss(tbase + 128, cdesc_a, cdesc_b, id_c128, 0);                   // dP = kMagic
#pragma unroll
    for (int k = 0; k < D / 32; ++k)
      umma_i8_ss(tbase + 128, umma_smem_desc(a_do + k * 32, 16, kSbo, kLay), umma_smem_desc(a_v + k * 32, 16, kSbo, kLay), id_s, 1);
    umma_commit(&sd_full);
  };
  auto issue_dv_dk = [&](int st, int dsb) {                        // dV = P^T dO, dK = dS^T Q (contraction over query rows)
    const uint32_t a_q = smem_u32(smem + L::off_q + st * L::kTile), a_do = smem_u32(smem + L::off_do + st * L::kTile);
    const uint32_t a_ds = a_ds0 + dsb * (128 * 128);
    umma_f16_ss(tbase + 256, cdesc_a, cdesc_b, id_cD, 0);                     // dV, dK partials = kMagic
    umma_f16_ss(tbase + 384, cdesc_a, cdesc_b, id_cD, 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      umma_i8_ss(tbase + 256, umma_smem_desc(a_p + k * 4096, 16, 1024, kSwz128), umma_smem_desc(a_do + k * 32 * D, 16, kSbo, kLay), id_t, 1);
      umma_i8_ss(tbase + 384, umma_smem_desc(a_ds + k * 4096, 16, 1024, kSwz128), umma_smem_desc(a_q + k * 32 * D, 16, kSbo, kLay), id_t, 1);
    }
    umma_commit(&parts_full);
  };
  auto issue_dq = [&](int dsb) {                                   // dQ = dS K (contraction over keys) -> cols 256..
    const uint32_t a_ds = a_ds0 + dsb * (128 * 128);
    umma_f16_ss(tbase + 256, cdesc_a, cdesc_b, id_cD, 0);                     // dQ partial = kMagic
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_i8_ss(tbase + 256, umma_smem_desc(a_ds + k * 32, 16, 1024, kSwz128), umma_smem_desc(a_k + k * 32 * D, 16, kSbo, kLay), id_q, 1);
    umma_commit(&dq_full);
  };

  if (leader) {
    mbar_expect_tx(&kv_full, 2 * L::kTile);
    tma_load_2d(smem + L::off_k, &tm_k, &kv_full, 0, (int)head_row0 + j * 128);
    tma_load_2d(smem + L::off_v, &tm_v, &kv_full, 0, (int)head_row0 + j * 128);
    load_qdo(t0, 0);
    mbar_wait(&kv_full, 0);
    mbar_wait(&qdo_full[0], 0);
    issue_s(0);
    issue_dp(0);
  }

  const uint32_t smem_base = smem_u32(smem);
  const int half = warp >> 2;                                   // column group handled by this thread
  const int row = (warp & 3) * 32 + lane;                       // TMEM lane: query row (S, dP, dQ) or key (dV, dK)
  const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  const float sk_f = __half2float(p.sk[head_row0 / 128 + j]);
  const float sv_f = __half2float(p.sv[head_row0 / 128 + j]);
  float2 dv_acc[DH / 2], dk_acc[DH / 2];                        // fp32x2: FFMA2 / FMUL2 / FADD2 halve the issue slots
#pragma unroll
  for (int d = 0; d < DH / 2; ++d) { dv_acc[d] = make_float2(0.f, 0.f); dk_acc[d] = make_float2(0.f, 0.f); }
  float c_dv_prev = 0.f, c_dk_prev = 0.f, c_dq_prev = 0.f;

  const float2 nM2 = make_float2(-kMagic, -kMagic);
  auto drain_dv_dk = [&](float c_dv, float c_dk) {
#pragma unroll
    for (int ch = 0; ch < DH / 16; ++ch) {
      uint32_t r[16], r2[16];
      tmem_ld16(lane_addr + 256 + half * DH + ch * 16, r);
      tmem_ld16(lane_addr + 384 + half * DH + ch * 16, r2);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        dv_acc[ch * 8 + c] = __ffma2_rn(__fadd2_rn(make_float2(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1])), nM2),
                                        make_float2(c_dv, c_dv), dv_acc[ch * 8 + c]);
        dk_acc[ch * 8 + c] = __ffma2_rn(__fadd2_rn(make_float2(__uint_as_float(r2[2 * c]), __uint_as_float(r2[2 * c + 1])), nM2),
                                        make_float2(c_dk, c_dk), dk_acc[ch * 8 + c]);
      }
    }
  };
  auto drain_dq = [&](float c_dq) {                 // dQ partial -> fp32 staging (swizzled 32-float atoms)
#pragma unroll
    for (int ch = 0; ch < DH / 16; ++ch) {
      uint32_t r[16];
      tmem_ld16(lane_addr + 256 + half * DH + ch * 16, r);
      tmem_ld_wait();
      const int col = half * DH + ch * 16;                         // first output column of this chunk
      const uint32_t atom = smem_base + L::off_dq + (col >> 5) * (128 * 128);
#pragma unroll
      for (int c = 0; c < 16; c += 4) {
        const float2 cq2 = make_float2(c_dq, c_dq), nbq2 = make_float2(-kMagic * c_dq, -kMagic * c_dq);   // c_dq: 22 significant bits
        const float2 o01 = __ffma2_rn(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), cq2, nbq2);
        const float2 o23 = __ffma2_rn(make_float2(__uint_as_float(r[c + 2]), __uint_as_float(r[c + 3])), cq2, nbq2);
        sts128f(atom + swz128(row, ((col & 31) + c) * 4), o01.x, o01.y, o23.x, o23.y);
      }
    }
  };
  auto reduce_dq = [&](int tile) {                                // dQ[tile] += staging (L2 reduction; k-tile order not fixed)
#pragma unroll
    for (int a = 0; a < D / 32; ++a)
      tma_reduce_add_2d(&tm_dq, smem + L::off_dq + a * (128 * 128), a * 32, (int)head_row0 + tile * 128);
    tma_store_commit();
  };

  constexpr float kPs = 1.0f / 1024.0f;                          // P is carried between the passes as fp16(1024 * P)
  for (int t = 0; t < nt; ++t) {
    const uint32_t ph = t & 1;
    const int tq = t0 + t;
    const size_t qrow = head_row0 + (size_t)tq * 128 + row;
    const float sq_f = __half2float(p.sq[head_row0 / 128 + tq]);
    const float sdo_f = __half2float(p.s_do[head_row0 / 128 + tq]);
    const float lse = (tq * 128 + row < p.S_valid) ? p.lse[qrow] : INFINITY;     // padded query rows of a ragged sequence: P = 0
    const float dlt = p.delta[qrow];
    const float c_s = magic_scale(sq_f * sk_f * p.qk_scale);
    const float c_dp = sdo_f * sv_f;
    QA_TLB(0);
    if (leader && t + 1 < nt) {                                    // next Q / dO tile: its stage was last read by dV/dK of t-1
      if (t > 0) mbar_wait(&parts_full, (t - 1) & 1);
      load_qdo(tq + 1, (t + 1) & 1);
    }
    mbar_wait(&sd_full, ph);
    tc_fence_after();
    QA_TLB(1);
    // ---- pass 1: P = exp2(fp16 logit - lse), kept as packed fp16 of 1024 * P (11-bit mantissa, no subnormals down to
    //      P = 6e-8) so that pass 2 needs no second exp2; tile amax of P and |dS|, row sum of dS
    __half2 pk[CW / 2];
    float amax_p = 0.f, amax_ds = 0.f;
    __half2 amax_ph = __float2half2_rn(0.f);
    float2 rs2acc = make_float2(0.f, 0.f);
    // magic accumulators: TMEM holds kMagic + x as a float; the scales carry 22 significant bits, so kMagic * scale is
    // exact and (kMagic + x) * scale - kMagic * scale is one rounding of x * scale
    const float nlse = 10.0f - lse, cdpk = magic_scale(c_dp * kPs);
    const float2 cs2 = make_float2(c_s, c_s), nbs2 = make_float2(-kMagic * c_s, -kMagic * c_s);
    const float2 cdp2 = make_float2(cdpk, cdpk), ndlt2 = make_float2(-dlt * kPs - kMagic * cdpk, -dlt * kPs - kMagic * cdpk);
    auto pass1 = [&](auto masked, auto tail) {                     // masked: the diagonal tile of a causal head; tail: ragged last k-tile
#pragma unroll
    for (int ch = 0; ch < CW / 16; ++ch) {
      uint32_t r[16], r2[16];
      tmem_ld16(lane_addr + half * CW + ch * 16, r);
      tmem_ld16(lane_addr + 128 + half * CW + ch * 16, r2);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const __half2 h = __float22half2_rn(__ffma2_rn(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), cs2, nbs2));
        const uint32_t hu = *reinterpret_cast<const uint32_t*>(&h);
        const float2 e = make_float2(fhadd_lo(hu, nlse), fhadd_hi(hu, nlse));           // float(logit) - lse + 10: one FHADD each
        float2 pp = make_float2(ex2_approx(e.x), ex2_approx(e.y));                     // 1024 * P
        if (decltype(masked)::value) {                                 // strict causal: keep key < query (same tile: col < row)
          const int col = half * CW + ch * 16 + c;
          if (col >= row) pp.x = 0.f;
          if (col + 1 >= row) pp.y = 0.f;
        }
        if (decltype(tail)::value) {                                   // padding keys of a ragged sequence
          const int col = half * CW + ch * 16 + c;
          if (col >= ktail) pp.x = 0.f;
          if (col + 1 >= ktail) pp.y = 0.f;
        }
        const __half2 pr = __float22half2_rn(pp);
        pk[ch * 8 + c / 2] = pr;
        amax_ph = __hmax2(amax_ph, pr);
        const float2 d = __fmul2_rn(pp, __ffma2_rn(make_float2(__uint_as_float(r2[c]), __uint_as_float(r2[c + 1])), cdp2, ndlt2));
        amax_ds = fmaxf(amax_ds, fmaxf(fabsf(d.x), fabsf(d.y)));
        rs2acc = __fadd2_rn(rs2acc, d);
      }
    }
    };
    if (CAUSAL && t == 0) pass1(std::true_type{}, std::false_type{});
    else if (!CAUSAL && ktail < 128) pass1(std::false_type{}, std::true_type{});
    else pass1(std::false_type{}, std::false_type{});
    amax_p = fmaxf(__low2float(amax_ph), __high2float(amax_ph));                       // of the rounded 1024 * P
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      amax_p = fmaxf(amax_p, __shfl_xor_sync(0xffffffffu, amax_p, o));
      amax_ds = fmaxf(amax_ds, __shfl_xor_sync(0xffffffffu, amax_ds, o));
    }
    if (lane == 0) { red_p[ph][warp] = amax_p; red_ds[ph][warp] = amax_ds; }
    rowsum_ds[ph][half][row] = rs2acc.x + rs2acc.y;
    QA_TLB(2);
    // ---- drain dV / dK of the previous tile (its MMAs ran while pass 1 executed)
    if (t > 0) {
      mbar_wait(&parts_full, (t - 1) & 1);
      tc_fence_after();
      QA_TLB(3);
      drain_dv_dk(c_dv_prev, c_dk_prev);
      QA_TLB(4);
    }
    tc_fence_before();
    if (leader) tma_store_wait_read();                             // the dQ staging tile may be rewritten after this barrier
    named_bar_sync(1, NT);                                        // amax partials visible; dV/dK partial columns drained
    QA_TLB(5);
    if (leader) {
      tc_fence_after();
      if (t > 0) issue_dq((t - 1) & 1);
      if (t + 1 < nt) {                                            // S of the next tile: everybody is past pass 1
        mbar_wait(&qdo_full[(t + 1) & 1], ((t + 1) >> 1) & 1);
        issue_s((t + 1) & 1);
      }
    }
    QA_TLB(6);
    // ---- tile-wide amax of P and |dS| (per-[Bq,Bkv]-tile quantisation, attention_int8.py:363-365, 403-405)
    amax_p = red_p[ph][0]; amax_ds = red_ds[ph][0];
#pragma unroll
    for (int w = 1; w < NW; ++w) { amax_p = fmaxf(amax_p, red_p[ph][w]); amax_ds = fmaxf(amax_ds, red_ds[ph][w]); }
    float rs_row = 0.f;                                                      // full-row sum of dS (query row = lane)
#pragma unroll
    for (int g = 0; g < NG; ++g) rs_row += rowsum_ds[ph][g][row];
    // K-smoothing term of dQ (LEDGER I-1): sm_scale * rowsum(dS) * k_mean is rank one per query row, so only the
    // row sums are accumulated here (one fp32 reduction per row and tile); qa_int8_bwd_finalize applies k_mean once.
    if (half == 0 && p.rowsum != nullptr) atomicAdd(p.rowsum + qrow, rs_row);
    const float sP = amax_p * (kPs / 127.0f), sdS = amax_ds * (1.0f / 127.0f);
    const float inv_p = amax_p > 0.f ? __fdividef(127.0f, amax_p) : 0.f;
    const float inv_ds = amax_ds > 0.f ? __fdividef(127.0f, amax_ds) : 0.f;
    // ---- pass 2: recompute P / dS from the packed logits, quantise (truncate toward zero), store both tiles as
    //      [q row][128 key bytes], 128B-swizzled (A operands of dV / dK (transposed) and dQ); dS is double-buffered
    const uint32_t ds_tile = smem_base + L::off_ds + ph * (128 * 128);
    const float cdpi = magic_scale(c_dp * kPs * inv_ds);
    const float2 cdpi2 = make_float2(cdpi, cdpi), ndlti2 = make_float2(-dlt * kPs * inv_ds - kMagic * cdpi, -dlt * kPs * inv_ds - kMagic * cdpi);
    const float2 invp2 = make_float2(inv_p, inv_p), magic2 = make_float2(8388608.0f, 8388608.0f);
    {                                                              // RN: the rounding mode is an instruction modifier
#pragma unroll
      for (int ch = 0; ch < CW / 16; ++ch) {
        uint32_t r2[16];
        tmem_ld16(lane_addr + 128 + half * CW + ch * 16, r2);
        tmem_ld_wait();
        uint32_t wp[4], wd[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t bp[4], bd[4];
#pragma unroll
          for (int e = 0; e < 4; e += 2) {
            const int c = q4 * 4 + e;
            const float2 pp = __half22float2(pk[ch * 8 + c / 2]);                  // 1024 * P from pass 1
            // dS / s_dS = P * (dP*c_dp - delta) / s_dS with 1/s_dS (and the 1/1024) folded into the FFMA2 constants
            const float2 dq = __fmul2_rn(pp, __ffma2_rn(make_float2(__uint_as_float(r2[c]), __uint_as_float(r2[c + 1])), cdpi2, ndlti2));
            // P >= 0: low byte of the biased sum = trunc(P / sP), or its nearest-even rounding in accuracy mode
            const float2 pq = RN ? __ffma2_rn(pp, invp2, magic2) : __ffma2_rz(pp, invp2, magic2);
            bp[e] = __float_as_uint(pq.x);
            bp[e + 1] = __float_as_uint(pq.y);
            bd[e] = (uint32_t)(RN ? __float2int_rn(dq.x) : __float2int_rz(dq.x));
            bd[e + 1] = (uint32_t)(RN ? __float2int_rn(dq.y) : __float2int_rz(dq.y));
          }
          wp[q4] = pack_low_bytes(bp[0], bp[1], bp[2], bp[3]);
          wd[q4] = pack_sat_s8x4((int)bd[0], (int)bd[1], (int)bd[2], (int)bd[3]);
        }
        const uint32_t off = swz128(row, half * CW + ch * 16);
        sts128(smem_base + L::off_p + off, wp[0], wp[1], wp[2], wp[3]);
        sts128(ds_tile + off, wd[0], wd[1], wd[2], wd[3]);
      }
    }
    QA_TLB(7);
    // ---- drain dQ of the previous tile (its MMA ran while pass 2 executed)
    if (t > 0) {
      mbar_wait(&dq_full, (t - 1) & 1);
      tc_fence_after();
      QA_TLB(8);
      drain_dq(c_dq_prev);
    }
    QA_TLB(9);
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(2, NT);                                        // P / dS tiles and the dQ staging tile are complete
    QA_TLB(10);
    if (leader) {
      tc_fence_after();
      if (t > 0) reduce_dq(tq - 1);
      if (t + 1 < nt) issue_dp((t + 1) & 1);
      QA_TLB(11);
      issue_dv_dk(t & 1, ph);
    }
    QA_TLB(12);
    c_dv_prev = sdo_f * sP;
    c_dk_prev = sdS * sq_f * p.sm_scale;
    c_dq_prev = magic_scale(sdS * sk_f * p.sm_scale);
  }
  // ---- pipeline tail: last tile's dV / dK / dQ
  mbar_wait(&parts_full, (nt - 1) & 1);
  tc_fence_after();
  drain_dv_dk(c_dv_prev, c_dk_prev);
  tc_fence_before();
  if (leader) tma_store_wait_read();
  named_bar_sync(1, NT);
  if (leader) { tc_fence_after(); issue_dq((nt - 1) & 1); }
  mbar_wait(&dq_full, (nt - 1) & 1);
  tc_fence_after();
  drain_dq(c_dq_prev);
  fence_proxy_async_smem();
  tc_fence_before();
  named_bar_sync(2, NT);
  if (leader) { reduce_dq(nq - 1); tma_store_wait_all(); }   // the last query tile, causal or not
  // ---- epilogue: dK_j, dV_j rows (row = key) in fp16
  const size_t krow = head_row0 + (size_t)j * 128 + row;
  __half* dk_dst = p.dk + krow * D + half * DH;
  __half* dv_dst = p.dv + krow * D + half * DH;
#pragma unroll
  for (int d = 0; d < DH; d += 8) {
    uint4 a, b;
    __half2 t2;
    t2 = __float22half2_rn(dk_acc[d / 2]); a.x = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dk_acc[d / 2 + 1]); a.y = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dk_acc[d / 2 + 2]); a.z = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dk_acc[d / 2 + 3]); a.w = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dv_acc[d / 2]); b.x = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dv_acc[d / 2 + 1]); b.y = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dv_acc[d / 2 + 2]); b.z = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dv_acc[d / 2 + 3]); b.w = *reinterpret_cast<uint32_t*>(&t2);
    *reinterpret_cast<uint4*>(dk_dst + d) = a;
    *reinterpret_cast<uint4*>(dv_dst + d) = b;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tbase);
}


// ---------------------------------------------------------------------------------------------------------
// Warp-specialised variant (default): 16 warps.  Warps 0..7 ("quantise" role) run pass 1, the tile-wide amax exchange
// and pass 2 of tile t; warps 8..15 ("drain" role) hold the fp32 dV / dK accumulators, drain the int32 partial
// products of tile t-1 and stage dQ; thread 256 issues TMA and tcgen05.mma.  setmaxnreg splits the register file
// 96 / 160 per thread.  Same TMEM / shared-memory plan, numerics and software pipeline as int8_bwd_kernel.
// Cross-role hand-over: sd_full (MMA -> quantise), pds_full (quantise -> leader: P / dS tiles stored, S / dP columns
// free), parts_full / dq_full (MMA -> drain), the tile scales through sc_ring (written before pds_full, read after
// parts_full of the same tile).  The quantise role never waits on dq_full (its parity could alias): completion of
// dQ(t-2), the last reader of the dS buffer it overwrites, is implied by sd_full(t), which was committed after it.
// ---------------------------------------------------------------------------------------------------------
template <int D, bool RN, bool CAUSAL>
__global__ void __launch_bounds__(512, 1)
int8_bwd_ws_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_dq, Int8BwdParams p) {
  using L = Int8BwdSmem<D>;
  constexpr int NG = 2;
  constexpr int DH = D / NG;                                    // output columns per drain thread
  constexpr int CW = 128 / NG;                                  // S / dP columns per quantise thread
  constexpr int NT = 256;                                       // threads per role
  constexpr uint32_t kLay = (D == 128) ? kSwz128 : kSwz64;      // operand rows of D bytes
  constexpr uint32_t kSbo = (D == 128) ? 1024 : 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t kv_full, qdo_full[2], sd_full, parts_full, dq_full, pds_full, s_free;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red_p[2][8], red_ds[2][8];
  __shared__ float rowsum_ds[2][NG][128];
  __shared__ float sc_ring[2][4];                               // per tile parity: c_dv, c_dk, c_dq

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool quant_role = warp < 8;
  const int rw = warp & 7;                                      // warp index inside the role
  // Three issuing threads in the drain role, on different SM sub-partitions (a tcgen05.mma / TMA issue blocks its thread):
  // 256: Q/dO loads, dQ, dV/dK;  288: S of the next tile;  320: the TMA reduce-add of the dQ staging tile.  dP of the next tile is
  // issued by thread 0 (quantise role) as soon as that role has released the dP columns
  const bool leader = (tid == 256), leader_sdp = (tid == 288), leader_red = (tid == 320);
  const int bh = blockIdx.y, j = blockIdx.x;
  const int nq = (p.S_valid + 127) / 128;                       // fully padded query tiles are skipped
  const int ktail = p.S_valid - j * 128;                        // valid keys of this k-tile (>= 128 unless it is the ragged last one)
  const int t0 = CAUSAL ? j : 0;                                // causal: k-tile j meets the query tiles j .. nq-1
  const int nt = nq - t0;
  const size_t head_row0 = (size_t)bh * p.S;

  if (tid == 0) {
    mbar_init(&kv_full, 1); mbar_init(&qdo_full[0], 1); mbar_init(&qdo_full[1], 1);
    mbar_init(&sd_full, 2);                                // the S commit (drain role) + the dP commit (quantise role)
    mbar_init(&parts_full, 1); mbar_init(&dq_full, 1); mbar_init(&pds_full, 8); mbar_init(&s_free, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  if (tid >= 64 && tid < 192) {                                 // constant atoms of the accumulator-initialising MMA
    const uint32_t v2 = tid < 128 ? kMagicElemA2 : kMagicElemB2;
    sts128(smem_u32(smem + L::off_c) + (tid - 64) * 16, v2, v2, v2, v2);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t smem_base = smem_u32(smem);
  const int half = rw >> 2;                                     // column group handled by this thread
  const int row = (rw & 3) * 32 + lane;                         // TMEM lane: query row (S, dP, dQ) or key (dV, dK)
  const uint32_t lane_addr = tbase + ((uint32_t)((rw & 3) * 32) << 16);
  const float sk_f = __half2float(p.sk[head_row0 / 128 + j]);
  const float sv_f = __half2float(p.sv[head_row0 / 128 + j]);
  constexpr float kPs = 1.0f / 1024.0f;                          // P is carried between the passes as fp16(1024 * P)

  if (quant_role) {
    // =========================== quantise role: pass 1, amax, pass 2 ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    for (int t = 0; t < nt; ++t) {
      const uint32_t ph = t & 1;
      const int tq = t0 + t;
      const size_t qrow = head_row0 + (size_t)tq * 128 + row;
      const float sq_f = __half2float(p.sq[head_row0 / 128 + tq]);
      const float sdo_f = __half2float(p.s_do[head_row0 / 128 + tq]);
      const float lse = (tq * 128 + row < p.S_valid) ? p.lse[qrow] : INFINITY;   // padded query rows of a ragged sequence: P = 0
      const float dlt = p.delta[qrow];
      const float c_s = magic_scale(sq_f * sk_f * p.qk_scale);
      const float c_dp = sdo_f * sv_f;
      if (warp == 0) QA_TLW(0, 0);
      mbar_wait(&sd_full, ph);
      tc_fence_after();
      if (warp == 0) QA_TLW(0, 1);
      // ---- pass 1: P = exp2(fp16 logit - lse), kept as packed fp16 of 1024 * P (11-bit mantissa, no subnormals down to
      //      P = 6e-8) so that pass 2 needs no second exp2; tile amax of P and |dS|, row sum of dS
      __half2 pk[CW / 2];
      float amax_p = 0.f, amax_ds = 0.f;
      __half2 amax_ph = __float2half2_rn(0.f);
      float2 rs2acc = make_float2(0.f, 0.f);
      // magic accumulators: TMEM holds kMagic + x as a float; the scales carry 22 significant bits, so kMagic * scale is
      // exact and (kMagic + x) * scale - kMagic * scale is one rounding of x * scale
      const float nlse = 10.0f - lse, cdpk = magic_scale(c_dp * kPs);
      const float2 cs2 = make_float2(c_s, c_s), nbs2 = make_float2(-kMagic * c_s, -kMagic * c_s);
      const float2 cdp2 = make_float2(cdpk, cdpk), ndlt2 = make_float2(-dlt * kPs - kMagic * cdpk, -dlt * kPs - kMagic * cdpk);
      auto pass1 = [&](auto masked, auto tail) {                     // masked: the diagonal tile of a causal head; tail: ragged last k-tile
  #pragma unroll
      for (int ch = 0; ch < CW / 16; ++ch) {
        uint32_t r[16], r2[16];
        tmem_ld16(lane_addr + half * CW + ch * 16, r);
        tmem_ld16(lane_addr + 128 + half * CW + ch * 16, r2);
        tmem_ld_wait();
  #pragma unroll
        for (int c = 0; c < 16; c += 2) {
          const __half2 h = __float22half2_rn(__ffma2_rn(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), cs2, nbs2));
          const uint32_t hu = *reinterpret_cast<const uint32_t*>(&h);
          const float2 e = make_float2(fhadd_lo(hu, nlse), fhadd_hi(hu, nlse));           // float(logit) - lse + 10: one FHADD each
          float2 pp = make_float2(ex2_approx(e.x), ex2_approx(e.y));                     // 1024 * P
          if (decltype(masked)::value) {                                 // strict causal: keep key < query (same tile: col < row)
            const int col = half * CW + ch * 16 + c;
            if (col >= row) pp.x = 0.f;
            if (col + 1 >= row) pp.y = 0.f;
          }
          if (decltype(tail)::value) {                                   // padding keys of a ragged sequence
            const int col = half * CW + ch * 16 + c;
            if (col >= ktail) pp.x = 0.f;
            if (col + 1 >= ktail) pp.y = 0.f;
          }
          const __half2 pr = __float22half2_rn(pp);
          pk[ch * 8 + c / 2] = pr;
          amax_ph = __hmax2(amax_ph, pr);
          const float2 d = __fmul2_rn(pp, __ffma2_rn(make_float2(__uint_as_float(r2[c]), __uint_as_float(r2[c + 1])), cdp2, ndlt2));
          amax_ds = fmaxf(amax_ds, fmaxf(fabsf(d.x), fabsf(d.y)));
          rs2acc = __fadd2_rn(rs2acc, d);
        }
      }
      };
      if (CAUSAL && t == 0) pass1(std::true_type{}, std::false_type{});
    else if (!CAUSAL && ktail < 128) pass1(std::false_type{}, std::true_type{});
    else pass1(std::false_type{}, std::false_type{});
      amax_p = fmaxf(__low2float(amax_ph), __high2float(amax_ph));                       // of the rounded 1024 * P
  #pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        amax_p = fmaxf(amax_p, __shfl_xor_sync(0xffffffffu, amax_p, o));
        amax_ds = fmaxf(amax_ds, __shfl_xor_sync(0xffffffffu, amax_ds, o));
      }
      if (lane == 0) { red_p[ph][rw] = amax_p; red_ds[ph][rw] = amax_ds; }
      rowsum_ds[ph][half][row] = rs2acc.x + rs2acc.y;

      if (warp == 0) QA_TLW(0, 2);
      named_bar_sync(3, NT);                                       // amax partials of the 8 quantise warps visible
      if (warp == 0) QA_TLW(0, 3);
      if (tid == 0) { tc_fence_before(); mbar_arrive(&s_free); }   // every quantise warp is past pass 1: the S columns are free
      // ---- tile-wide amax of P and |dS| (per-[Bq,Bkv]-tile quantisation, attention_int8.py:363-365, 403-405)
      amax_p = red_p[ph][0]; amax_ds = red_ds[ph][0];
  #pragma unroll
      for (int w = 1; w < 8; ++w) { amax_p = fmaxf(amax_p, red_p[ph][w]); amax_ds = fmaxf(amax_ds, red_ds[ph][w]); }
      float rs_row = 0.f;                                                      // full-row sum of dS (query row = lane)
  #pragma unroll
      for (int g = 0; g < 2; ++g) rs_row += rowsum_ds[ph][g][row];
      // K-smoothing term of dQ (LEDGER I-1): sm_scale * rowsum(dS) * k_mean is rank one per query row, so only the
      // row sums are accumulated here (one fp32 reduction per row and tile); qa_int8_bwd_finalize applies k_mean once.
      if (half == 0 && p.rowsum != nullptr) atomicAdd(p.rowsum + qrow, rs_row);
      const float sP = amax_p * (kPs / 127.0f), sdS = amax_ds * (1.0f / 127.0f);
      const float inv_p = amax_p > 0.f ? __fdividef(127.0f, amax_p) : 0.f;
      const float inv_ds = amax_ds > 0.f ? __fdividef(127.0f, amax_ds) : 0.f;
      if (tid == 0) {                                              // tile scales for the drain role (read after parts_full(t))
        sc_ring[ph][0] = sdo_f * sP;
        sc_ring[ph][1] = sdS * sq_f * p.sm_scale;
        sc_ring[ph][2] = magic_scale(sdS * sk_f * p.sm_scale);
      }
      // the P tile and this dS buffer were last read by dV/dK of tile t-1 (dS[ph] also by dQ of t-2, which completed
      // before sd_full(t) was signalled)
      if (t > 0) mbar_wait(&parts_full, (t - 1) & 1);
      if (warp == 0) QA_TLW(0, 4);
      // ---- pass 2: recompute P / dS from the packed logits, quantise (truncate toward zero), store both tiles as
      //      [q row][128 key bytes], 128B-swizzled (A operands of dV / dK (transposed) and dQ); dS is double-buffered
      const uint32_t ds_tile = smem_base + L::off_ds + ph * (128 * 128);
      const float cdpi = magic_scale(c_dp * kPs * inv_ds);
      const float2 cdpi2 = make_float2(cdpi, cdpi), ndlti2 = make_float2(-dlt * kPs * inv_ds - kMagic * cdpi, -dlt * kPs * inv_ds - kMagic * cdpi);
      const float2 invp2 = make_float2(inv_p, inv_p), magic2 = make_float2(8388608.0f, 8388608.0f);
      {                                                              // RN: the rounding mode is an instruction modifier
  #pragma unroll
        for (int ch = 0; ch < CW / 16; ++ch) {
          uint32_t r2[16];
          tmem_ld16(lane_addr + 128 + half * CW + ch * 16, r2);
          tmem_ld_wait();
          uint32_t wp[4], wd[4];
  #pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t bp[4], bd[4];
  #pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const int c = q4 * 4 + e;
              const float2 pp = __half22float2(pk[ch * 8 + c / 2]);                  // 1024 * P from pass 1
              // dS / s_dS = P * (dP*c_dp - delta) / s_dS with 1/s_dS (and the 1/1024) folded into the FFMA2 constants
              const float2 dq = __fmul2_rn(pp, __ffma2_rn(make_float2(__uint_as_float(r2[c]), __uint_as_float(r2[c + 1])), cdpi2, ndlti2));
              // P >= 0: low byte of the biased sum = trunc(P / sP), or its nearest-even rounding in accuracy mode
              const float2 pq = RN ? __ffma2_rn(pp, invp2, magic2) : __ffma2_rz(pp, invp2, magic2);
              bp[e] = __float_as_uint(pq.x);
              bp[e + 1] = __float_as_uint(pq.y);
              bd[e] = (uint32_t)(RN ? __float2int_rn(dq.x) : __float2int_rz(dq.x));
              bd[e + 1] = (uint32_t)(RN ? __float2int_rn(dq.y) : __float2int_rz(dq.y));
            }
            wp[q4] = pack_low_bytes(bp[0], bp[1], bp[2], bp[3]);
            wd[q4] = pack_sat_s8x4((int)bd[0], (int)bd[1], (int)bd[2], (int)bd[3]);
          }
          const uint32_t off = swz128(row, half * CW + ch * 16);
          sts128(smem_base + L::off_p + off, wp[0], wp[1], wp[2], wp[3]);
          sts128(ds_tile + off, wd[0], wd[1], wd[2], wd[3]);
        }
      }

      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full);
      if (warp == 0) QA_TLW(0, 5);
      QA_TLW(0, 6 + warp);                                         // pass-2 end of every quantise warp (skew between them)
      if (warp == 0) {
        // dP of the next tile is issued from the role that waits for it, as soon as every quantise warp has released the dP
        // columns (pds_full), instead of when the drain role comes round to it (a tcgen05.mma issue blocks its thread until
        // the tensor pipe takes the instruction: here it blocks a warp that has nothing else to do)
        if (lane == 0 && t + 1 < nt) {
          constexpr uint32_t id_s = umma_idesc(2, 1, 1, 0, 0, 128, 128), id_c128 = umma_idesc(1, 0, 0, 0, 0, 128, 128);
          const int st = (t + 1) & 1;
          mbar_wait(&pds_full, ph);
          mbar_wait(&qdo_full[st], ((t + 1) >> 1) & 1);
          tc_fence_after();
          const uint32_t a_do = smem_u32(smem + L::off_do + st * L::kTile), a_v = smem_u32(smem + L::off_v);
          umma_f16_ss(tbase + 128, umma_smem_desc(smem_u32(smem + L::off_c), 16, 0, kLay), umma_smem_desc(smem_u32(smem + L::off_c + 1024), 16, 0, kLay),
                      id_c128, 0);                                  // dP = kMagic
  #pragma unroll
          for (int k = 0; k < D / 32; ++k)
            umma_i8_ss(tbase + 128, umma_smem_desc(a_do + k * 32, 16, kSbo, kLay), umma_smem_desc(a_v + k * 32, 16, kSbo, kLay), id_s, 1);
          umma_commit(&sd_full);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== drain role: accumulators, dQ staging, TMA / MMA issue ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
    constexpr uint32_t id_s = umma_idesc(2, 1, 1, 0, 0, 128, 128);     // S, dP: A, B K-major, N = 128
    constexpr uint32_t id_t = umma_idesc(2, 1, 1, 1, 1, 128, D);       // dV, dK: A^T (MN-major), B MN-major, N = D
    constexpr uint32_t id_q = umma_idesc(2, 1, 1, 0, 1, 128, D);       // dQ: A K-major, B MN-major
    // "magic" accumulators (qa_ptx.cuh): a kind::f16 MMA over two constant atoms sets an accumulator to kMagic before the
    // kind::i8 MMAs add to it; all sixteen 8-row groups read the same 1 KB atom (SBO = 0)
    constexpr uint32_t id_c128 = umma_idesc(1, 0, 0, 0, 0, 128, 128), id_cD = umma_idesc(1, 0, 0, 0, 0, 128, D);
    const uint64_t cdesc_a = umma_smem_desc(smem_u32(smem + L::off_c), 16, 0, kLay), cdesc_b = umma_smem_desc(smem_u32(smem + L::off_c + 1024), 16, 0, kLay);
    const uint32_t a_k = smem_u32(smem + L::off_k), a_v = smem_u32(smem + L::off_v);
    const uint32_t a_p = smem_u32(smem + L::off_p), a_ds0 = smem_u32(smem + L::off_ds);
    auto load_qdo = [&](int tile, int st) {
      mbar_expect_tx(&qdo_full[st], 2 * L::kTile);
      tma_load_2d(smem + L::off_q + st * L::kTile, &tm_q, &qdo_full[st], 0, (int)head_row0 + tile * 128);
      tma_load_2d(smem + L::off_do + st * L::kTile, &tm_do, &qdo_full[st], 0, (int)head_row0 + tile * 128);
    };
    auto issue_s = [&](int st) {                                     // S = Q K^T (issued one barrier ahead of dP: the S columns
      const uint32_t a_q = smem_u32(smem + L::off_q + st * L::kTile);          // are free as soon as pass 1 is over)
      umma_f16_ss(tbase + 0, cdesc_a, cdesc_b, id_c128, 0);                     // S = kMagic (qa_ptx.cuh)
  #pragma unroll
      for (int k = 0; k < D / 32; ++k)
        umma_i8_ss(tbase + 0, umma_smem_desc(a_q + k * 32, 16, kSbo, kLay), umma_smem_desc(a_k + k * 32, 16, kSbo, kLay), id_s, 1);
    };
    auto issue_dp = [&](int st) {                                    // dP = dO V^T; the commit also covers the earlier S MMAs
      const uint32_t a_do = smem_u32(smem + L::off_do + st * L::kTile);
      umma_f16_ss(tbase + 128, cdesc_a, cdesc_b, id_c128, 0);                   // dP = kMagic
  #pragma unroll
      for (int k = 0; k < D / 32; ++k)
        umma_i8_ss(tbase + 128, umma_smem_desc(a_do + k * 32, 16, kSbo, kLay), umma_smem_desc(a_v + k * 32, 16, kSbo, kLay), id_s, 1);
      umma_commit(&sd_full);
    };
    auto issue_dv_dk = [&](int st, int dsb) {                        // dV = P^T dO, dK = dS^T Q (contraction over query rows)
      const uint32_t a_q = smem_u32(smem + L::off_q + st * L::kTile), a_do = smem_u32(smem + L::off_do + st * L::kTile);
      const uint32_t a_ds = a_ds0 + dsb * (128 * 128);
      umma_f16_ss(tbase + 256, cdesc_a, cdesc_b, id_cD, 0);                     // dV, dK partials = kMagic
      umma_f16_ss(tbase + 384, cdesc_a, cdesc_b, id_cD, 0);
  #pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_i8_ss(tbase + 256, umma_smem_desc(a_p + k * 4096, 16, 1024, kSwz128), umma_smem_desc(a_do + k * 32 * D, 16, kSbo, kLay), id_t, 1);
        umma_i8_ss(tbase + 384, umma_smem_desc(a_ds + k * 4096, 16, 1024, kSwz128), umma_smem_desc(a_q + k * 32 * D, 16, kSbo, kLay), id_t, 1);
      }
      umma_commit(&parts_full);
    };
    auto issue_dq = [&](int dsb) {                                   // dQ = dS K (contraction over keys) -> cols 256..
      const uint32_t a_ds = a_ds0 + dsb * (128 * 128);
      umma_f16_ss(tbase + 256, cdesc_a, cdesc_b, id_cD, 0);                     // dQ partial = kMagic
  #pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_i8_ss(tbase + 256, umma_smem_desc(a_ds + k * 32, 16, 1024, kSwz128), umma_smem_desc(a_k + k * 32 * D, 16, kSbo, kLay), id_q, 1);
      umma_commit(&dq_full);
    };

    if (leader) {
      mbar_expect_tx(&kv_full, 2 * L::kTile);
      tma_load_2d(smem + L::off_k, &tm_k, &kv_full, 0, (int)head_row0 + j * 128);
      tma_load_2d(smem + L::off_v, &tm_v, &kv_full, 0, (int)head_row0 + j * 128);
      load_qdo(t0, 0);
      mbar_wait(&kv_full, 0);
      mbar_wait(&qdo_full[0], 0);
      issue_s(0);
      umma_commit(&sd_full);
      issue_dp(0);
    }
    float2 dv_acc[DH / 2], dk_acc[DH / 2];                        // fp32x2: FFMA2 / FMUL2 / FADD2 halve the issue slots
#pragma unroll
    for (int d = 0; d < DH / 2; ++d) { dv_acc[d] = make_float2(0.f, 0.f); dk_acc[d] = make_float2(0.f, 0.f); }
    const float2 nM2 = make_float2(-kMagic, -kMagic);
    auto drain_dv_dk = [&](float c_dv, float c_dk) {
  #pragma unroll
      for (int ch = 0; ch < DH / 16; ++ch) {
        uint32_t r[16], r2[16];
        tmem_ld16(lane_addr + 256 + half * DH + ch * 16, r);
        tmem_ld16(lane_addr + 384 + half * DH + ch * 16, r2);
        tmem_ld_wait();
  #pragma unroll
        for (int c = 0; c < 8; ++c) {
          dv_acc[ch * 8 + c] = __ffma2_rn(__fadd2_rn(make_float2(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1])), nM2),
                                          make_float2(c_dv, c_dv), dv_acc[ch * 8 + c]);
          dk_acc[ch * 8 + c] = __ffma2_rn(__fadd2_rn(make_float2(__uint_as_float(r2[2 * c]), __uint_as_float(r2[2 * c + 1])), nM2),
                                          make_float2(c_dk, c_dk), dk_acc[ch * 8 + c]);
        }
      }
    };
    auto drain_dq = [&](float c_dq) {                 // dQ partial -> fp32 staging (swizzled 32-float atoms)
  #pragma unroll
      for (int ch = 0; ch < DH / 16; ++ch) {
        uint32_t r[16];
        tmem_ld16(lane_addr + 256 + half * DH + ch * 16, r);
        tmem_ld_wait();
        const int col = half * DH + ch * 16;                         // first output column of this chunk
        const uint32_t atom = smem_base + L::off_dq + (col >> 5) * (128 * 128);
  #pragma unroll
        for (int c = 0; c < 16; c += 4) {
          const float2 cq2 = make_float2(c_dq, c_dq), nbq2 = make_float2(-kMagic * c_dq, -kMagic * c_dq);   // c_dq: 22 significant bits
          const float2 o01 = __ffma2_rn(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), cq2, nbq2);
          const float2 o23 = __ffma2_rn(make_float2(__uint_as_float(r[c + 2]), __uint_as_float(r[c + 3])), cq2, nbq2);
          sts128f(atom + swz128(row, ((col & 31) + c) * 4), o01.x, o01.y, o23.x, o23.y);
        }
      }
    };
    auto reduce_dq = [&](int tile) {                                // dQ[tile] += staging (L2 reduction; k-tile order not fixed)
  #pragma unroll
      for (int a = 0; a < D / 32; ++a)
        tma_reduce_add_2d(&tm_dq, smem + L::off_dq + a * (128 * 128), a * 32, (int)head_row0 + tile * 128);
      tma_store_commit();
    };

    for (int t = 0; t < nt; ++t) {
      const uint32_t ph = t & 1;
      const int tq = t0 + t;
      if (warp == 8) QA_TLW(1, 6);
      if (leader && t + 1 < nt) {                                  // next Q / dO tile: its stage was last read by dV/dK of t-1
        if (t > 0) mbar_wait(&parts_full, (t - 1) & 1);
        load_qdo(tq + 1, (t + 1) & 1);
      }
      float c_dq_prev = 0.f;
      if (t > 0) {                                                 // dV / dK partials of tile t-1
        mbar_wait(&parts_full, (t - 1) & 1);
        tc_fence_after();
        if (warp == 8) QA_TLW(1, 7);
        const float c_dv_prev = sc_ring[(t - 1) & 1][0], c_dk_prev = sc_ring[(t - 1) & 1][1];
        c_dq_prev = sc_ring[(t - 1) & 1][2];
        drain_dv_dk(c_dv_prev, c_dk_prev);
        if (warp == 8) QA_TLW(1, 8);
      }
      tc_fence_before();
      if (leader_red) tma_store_wait_read();                       // the dQ staging tile may be rewritten after this barrier
      named_bar_sync(1, NT);                                      // dV/dK partial columns drained by the whole role
      if (warp == 8) QA_TLW(1, 9);
      if (leader && t > 0) { tc_fence_after(); issue_dq((t - 1) & 1); }
      if (leader_sdp && t + 1 < nt) {                              // S of the next tile as soon as pass 1 of this one is over
        mbar_wait(&s_free, ph);
        tc_fence_after();
        mbar_wait(&qdo_full[(t + 1) & 1], ((t + 1) >> 1) & 1);
        issue_s((t + 1) & 1);
        umma_commit(&sd_full);                                     // second arrival: the dP commit of the quantise role
      }
      if (warp == 8) QA_TLW(1, 10);
      if (t > 0) {                                                 // dQ partial of tile t-1 -> staging tile
        mbar_wait(&dq_full, (t - 1) & 1);
        tc_fence_after();
        if (warp == 8) QA_TLW(1, 11);
        drain_dq(c_dq_prev);
        if (warp == 8) QA_TLW(1, 12);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(2, NT);                                      // staging tile complete, dQ partial columns drained
      if (warp == 8) QA_TLW(1, 13);
      if (leader_red && t > 0) reduce_dq(tq - 1);
      if (leader) {
        QA_TLW(1, 0);
        mbar_wait(&pds_full, ph);                                  // P / dS tiles of tile t stored
        QA_TLW(1, 1);
        tc_fence_after();
        QA_TLW(1, 14);
        issue_dv_dk(t & 1, ph);
        QA_TLW(1, 15);
      }
    }
    // ---- pipeline tail: last tile's dV / dK / dQ
    mbar_wait(&parts_full, (nt - 1) & 1);
    tc_fence_after();
    {
      const float c_dv_prev = sc_ring[(nt - 1) & 1][0], c_dk_prev = sc_ring[(nt - 1) & 1][1], c_dq_prev = sc_ring[(nt - 1) & 1][2];
      drain_dv_dk(c_dv_prev, c_dk_prev);
      tc_fence_before();
      if (leader_red) tma_store_wait_read();
      named_bar_sync(1, NT);
      if (leader) { tc_fence_after(); issue_dq((nt - 1) & 1); }
      mbar_wait(&dq_full, (nt - 1) & 1);
      tc_fence_after();
      drain_dq(c_dq_prev);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(2, NT);
    if (leader_red) { reduce_dq(nq - 1); tma_store_wait_all(); }
    // ---- epilogue: dK_j, dV_j rows (row = key) in fp16
    const size_t krow = head_row0 + (size_t)j * 128 + row;
    __half* dk_dst = p.dk + krow * D + half * DH;
    __half* dv_dst = p.dv + krow * D + half * DH;
  #pragma unroll
    for (int d = 0; d < DH; d += 8) {
      uint4 a, b;
      __half2 t2;
      t2 = __float22half2_rn(dk_acc[d / 2]); a.x = *reinterpret_cast<uint32_t*>(&t2);
      t2 = __float22half2_rn(dk_acc[d / 2 + 1]); a.y = *reinterpret_cast<uint32_t*>(&t2);
      t2 = __float22half2_rn(dk_acc[d / 2 + 2]); a.z = *reinterpret_cast<uint32_t*>(&t2);
      t2 = __float22half2_rn(dk_acc[d / 2 + 3]); a.w = *reinterpret_cast<uint32_t*>(&t2);
      t2 = __float22half2_rn(dv_acc[d / 2]); b.x = *reinterpret_cast<uint32_t*>(&t2);
      t2 = __float22half2_rn(dv_acc[d / 2 + 1]); b.y = *reinterpret_cast<uint32_t*>(&t2);
      t2 = __float22half2_rn(dv_acc[d / 2 + 2]); b.z = *reinterpret_cast<uint32_t*>(&t2);
      t2 = __float22half2_rn(dv_acc[d / 2 + 3]); b.w = *reinterpret_cast<uint32_t*>(&t2);
      *reinterpret_cast<uint4*>(dk_dst + d) = a;
      *reinterpret_cast<uint4*>(dv_dst + d) = b;
    }

  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tbase);
}

// ---------------------------------------------------------------------------------------------------------
// Backward for quantisation blocks smaller than the MMA tile: Bq, Bkv in {32, 64, 128} (the reference's tunables,
// PowerOfTwoFragment(32, 256, 32), attention_int8.py:155-158; 32/32 is its untuned default).  The 128 x 128 tile then
// holds (128/Bq) x (128/Bkv) quantisation blocks of P and dS, each with its own scale (attention_int8.py:363-365,
// 403-405), dO and Q have one scale per Bq rows, K and V one per Bkv keys.  S and dP scale per (row block, key block) and
// stay one MMA each; the three contractions whose scales change ALONG the reduction axis are issued one quantisation
// block at a time (dV, dK: over the query-row blocks; dQ: over the key blocks) - each a subset of the k-steps of the
// full-tile MMA into a freshly initialised accumulator - and drained with the scale of that block and of the thread's own
// row.  Correctness-first schedule: the phases of a tile pair run one after the other (no cross-tile software pipeline),
// 8 warps, thread = (row, column half).  Same numerics as int8_bwd_kernel otherwise (magic accumulators, fp16-carried P).
// ---------------------------------------------------------------------------------------------------------
// SAGE (SageBwd): dP = dO V^T is NOT quantised - SageAttention3 keeps this one contraction in fp16
// because dP - delta cancels catastrophically (the reference quantises it too, attention_int8.py:380-384).  tm_v / tm_do16
// then describe the fp16 V and dO tensors, the MMA is kind::f16 with fp32 accumulation and pass 1 / pass 2 read dP as a
// float; dV = P^T dO still uses the int8 dO.
template <int D, bool RN, bool SAGE>
__global__ void __launch_bounds__(256, 1)
int8_bwd_blk_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                    const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_do16, Int8BwdParams p,
                    int BQ, int BK) {
  using L = std::conditional_t<SAGE, Int8BwdSageSmem<D>, Int8BwdSmem<D>>;
  constexpr int kAtoms16 = D / 64, kAtom16 = 128 * 128;          // fp16 operand atoms ([128][64] x 2 B)
  constexpr int DH = D / 2, CW = 64, NT = 256;
  constexpr uint32_t kLay = (D == 128) ? kSwz128 : kSwz64;
  constexpr uint32_t kSbo = (D == 128) ? 1024 : 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t kv_full, qdo_full, mma_done;
  __shared__ uint32_t tmem_base_s;
  __shared__ uint32_t amax_p_s[2][4][4], amax_ds_s[2][4][4];     // float bits (values >= 0): atomicMax on the bit pattern
  __shared__ float rowsum_ds[2][128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool leader = (tid == 0);
  const int bh = blockIdx.y, j = blockIdx.x;
  const int nq = (p.S_valid + 127) / 128;                       // fully padded query tiles are skipped
  const int ktail = p.S_valid - j * 128;                        // valid keys of this k-tile (>= 128 unless it is the ragged last one)
  const int NRB = 128 / BQ, NCB = 128 / BK;
  const size_t head_row0 = (size_t)bh * p.S;

  if (leader) {
    mbar_init(&kv_full, 1); mbar_init(&qdo_full, 1); mbar_init(&mma_done, 1);
    fence_mbar_init();
  }
  if (tid < 32) { (&amax_p_s[0][0][0])[tid] = 0u; (&amax_ds_s[0][0][0])[tid] = 0u; }
  if (warp == 1) tmem_alloc<512>(&tmem_base_s);
  if (tid >= 64 && tid < 192) {                                 // constant atoms of the accumulator-initialising MMA
    const uint32_t v2 = tid < 128 ? kMagicElemA2 : kMagicElemB2;
    sts128(smem_u32(smem + L::off_c) + (tid - 64) * 16, v2, v2, v2, v2);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const uint32_t smem_base = smem_u32(smem);

  constexpr uint32_t id_s = umma_idesc(2, 1, 1, 0, 0, 128, 128);
  constexpr uint32_t id_t = umma_idesc(2, 1, 1, 1, 1, 128, D);
  constexpr uint32_t id_q = umma_idesc(2, 1, 1, 0, 1, 128, D);
  constexpr uint32_t id_c128 = umma_idesc(1, 0, 0, 0, 0, 128, 128), id_cD = umma_idesc(1, 0, 0, 0, 0, 128, D);
  const uint64_t cdesc_a = umma_smem_desc(smem_u32(smem + L::off_c), 16, 0, kLay), cdesc_b = umma_smem_desc(smem_u32(smem + L::off_c + 1024), 16, 0, kLay);
  const uint32_t a_k = smem_u32(smem + L::off_k), a_v = smem_u32(smem + L::off_v);
  const uint32_t a_q = smem_u32(smem + L::off_q), a_do = smem_u32(smem + L::off_do);
  const uint32_t a_p = smem_u32(smem + L::off_p), a_ds = smem_u32(smem + L::off_ds);

  if (leader) {
    if (SAGE) {
      mbar_expect_tx(&kv_full, L::kTile + 128 * D * 2);
      tma_load_2d(smem + L::off_k, &tm_k, &kv_full, 0, (int)head_row0 + j * 128);
#pragma unroll
      for (int a = 0; a < kAtoms16; ++a)                           // fp16 V tile, one [128][64] atom per 64 columns
        tma_load_2d(smem + L::off_c + 2048 + a * kAtom16, &tm_v, &kv_full, a * 64, (int)head_row0 + j * 128);
    } else {
      mbar_expect_tx(&kv_full, 2 * L::kTile);
      tma_load_2d(smem + L::off_k, &tm_k, &kv_full, 0, (int)head_row0 + j * 128);
      tma_load_2d(smem + L::off_v, &tm_v, &kv_full, 0, (int)head_row0 + j * 128);
    }
    mbar_wait(&kv_full, 0);
  }

  const int half = warp >> 2;
  const int row = (warp & 3) * 32 + lane;                       // query row (S, dP, dQ) or key (dV, dK)
  const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  const int rb = row / BQ;                                      // row block of this thread's QUERY row
  const int kb = row / BK;                                      // key block of this thread's KEY row
  float sk_c[4], sv_c[4];                                       // scales of the key blocks of this k-tile
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    sk_c[c] = c < NCB ? __half2float(p.sk[(head_row0 + (size_t)j * 128) / BK + c]) : 0.f;
    sv_c[c] = c < NCB ? __half2float(p.sv[(head_row0 + (size_t)j * 128) / BK + c]) : 0.f;
  }
  float2 dv_acc[DH / 2], dk_acc[DH / 2];
#pragma unroll
  for (int d = 0; d < DH / 2; ++d) { dv_acc[d] = make_float2(0.f, 0.f); dk_acc[d] = make_float2(0.f, 0.f); }
  const float2 nM2 = make_float2(-kMagic, -kMagic);
  constexpr float kPs = 1.0f / 1024.0f;
  uint32_t n_mma = 0;                                           // completed phases of mma_done (same count in every thread)

  for (int t = 0; t < nq; ++t) {
    const uint32_t ab = t & 1;                                   // amax buffer of this tile
    const int tq = t;
    const size_t qrow = head_row0 + (size_t)t * 128 + row;
    const size_t qblk0 = (head_row0 + (size_t)t * 128) / BQ;
    const float sq_f = __half2float(p.sq[qblk0 + rb]);
    const float sdo_f = __half2float(p.s_do[qblk0 + rb]);
    const float lse = (tq * 128 + row < p.S_valid) ? p.lse[qrow] : INFINITY;     // padded query rows of a ragged sequence: P = 0
    const float dlt = p.delta[qrow];
    if (leader) {
      mbar_expect_tx(&qdo_full, 2 * L::kTile + (SAGE ? 128 * D * 2 : 0));
      tma_load_2d(smem + L::off_q, &tm_q, &qdo_full, 0, (int)head_row0 + t * 128);
      tma_load_2d(smem + L::off_do, &tm_do, &qdo_full, 0, (int)head_row0 + t * 128);
      if (SAGE) {
#pragma unroll
        for (int a = 0; a < kAtoms16; ++a)
          tma_load_2d(smem + L::off_c + 2048 + 128 * D * 2 + a * kAtom16, &tm_do16, &qdo_full, a * 64, (int)head_row0 + t * 128);
      }
      mbar_wait(&qdo_full, t & 1);
      umma_f16_ss(tbase + 0, cdesc_a, cdesc_b, id_c128, 0);
      if (!SAGE) umma_f16_ss(tbase + 128, cdesc_a, cdesc_b, id_c128, 0);
#pragma unroll
      for (int k = 0; k < D / 32; ++k) {
        umma_i8_ss(tbase + 0, umma_smem_desc(a_q + k * 32, 16, kSbo, kLay), umma_smem_desc(a_k + k * 32, 16, kSbo, kLay), id_s, 1);
        if (!SAGE) umma_i8_ss(tbase + 128, umma_smem_desc(a_do + k * 32, 16, kSbo, kLay), umma_smem_desc(a_v + k * 32, 16, kSbo, kLay), id_s, 1);
      }
      if (SAGE) {                                                  // dP = dO V^T in fp16 (fp32 accumulation): 16 columns of D per MMA
        constexpr uint32_t id_dp16 = umma_idesc(1, 0, 0, 0, 0, 128, 128);
        const uint32_t a_v16 = smem_base + L::off_c + 2048, a_do16 = a_v16 + 128 * D * 2;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t o = (k >> 2) * kAtom16 + (k & 3) * 32;
          umma_f16_ss(tbase + 128, umma_smem_desc(a_do16 + o, 16, 1024, kSwz128), umma_smem_desc(a_v16 + o, 16, 1024, kSwz128), id_dp16, k > 0);
        }
      }
      umma_commit(&mma_done);
    }
    mbar_wait(&mma_done, n_mma & 1); ++n_mma;
    tc_fence_after();
    // ---- pass 1: P (carried as fp16 of 1024 * P), block amax of P and |dS|, row sum of dS
    __half2 pk[CW / 2];
    float am_p[2] = {0.f, 0.f}, am_ds[2] = {0.f, 0.f};            // this thread's (up to two) key blocks
    float2 rs2acc = make_float2(0.f, 0.f);
    const float nlse = 10.0f - lse;
    const int cb_first = (half * CW) / BK;
#pragma unroll
    for (int ch = 0; ch < CW / 16; ++ch) {
      const int cb = (half * CW + ch * 16) / BK;                 // a 16-column chunk lies inside one key block (Bkv >= 32)
      const float c_s = magic_scale(sq_f * sk_c[cb & 3] * p.qk_scale);
      const float cdpk = SAGE ? kPs : magic_scale(sdo_f * sv_c[cb & 3] * kPs);      // SAGE: dP is the fp32 accumulator itself
      const float2 cs2 = make_float2(c_s, c_s), nbs2 = make_float2(-kMagic * c_s, -kMagic * c_s);
      const float dpb = SAGE ? -dlt * kPs : -dlt * kPs - kMagic * cdpk;
      const float2 cdp2 = make_float2(cdpk, cdpk), ndlt2 = make_float2(dpb, dpb);
      uint32_t r[16], r2[16];
      tmem_ld16(lane_addr + half * CW + ch * 16, r);
      tmem_ld16(lane_addr + 128 + half * CW + ch * 16, r2);
      tmem_ld_wait();
      float mp = 0.f, md = 0.f;
#pragma unroll
      for (int c = 0; c < 16; c += 2) {
        const __half2 h = __float22half2_rn(__ffma2_rn(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), cs2, nbs2));
        const uint32_t hu = *reinterpret_cast<const uint32_t*>(&h);
        float2 pp = make_float2(ex2_approx(fhadd_lo(hu, nlse)), ex2_approx(fhadd_hi(hu, nlse)));     // 1024 * P
        if (ktail < 128) {                                           // padding keys of a ragged sequence
          const int col = half * CW + ch * 16 + c;
          if (col >= ktail) pp.x = 0.f;
          if (col + 1 >= ktail) pp.y = 0.f;
        }
        const __half2 pr = __float22half2_rn(pp);
        pk[ch * 8 + c / 2] = pr;
        mp = fmaxf(mp, fmaxf(__low2float(pr), __high2float(pr)));
        const float2 d = __fmul2_rn(pp, __ffma2_rn(make_float2(__uint_as_float(r2[c]), __uint_as_float(r2[c + 1])), cdp2, ndlt2));
        md = fmaxf(md, fmaxf(fabsf(d.x), fabsf(d.y)));
        rs2acc = __fadd2_rn(rs2acc, d);
      }
      am_p[cb - cb_first] = fmaxf(am_p[cb - cb_first], mp);
      am_ds[cb - cb_first] = fmaxf(am_ds[cb - cb_first], md);
    }
    // the 32 rows of a warp share one row block (Bq >= 32): warp maximum, then one shared-memory atomic per block
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        am_p[i] = fmaxf(am_p[i], __shfl_xor_sync(0xffffffffu, am_p[i], o));
        am_ds[i] = fmaxf(am_ds[i], __shfl_xor_sync(0xffffffffu, am_ds[i], o));
      }
      if (lane == 0 && cb_first + i < NCB && (i == 0 || BK == 32)) {
        atomicMax(&amax_p_s[ab][rb][cb_first + i], __float_as_uint(am_p[i]));
        atomicMax(&amax_ds_s[ab][rb][cb_first + i], __float_as_uint(am_ds[i]));
      }
    }
    rowsum_ds[half][row] = rs2acc.x + rs2acc.y;
    tc_fence_before();
    named_bar_sync(1, NT);                                       // block maxima complete; S / dP read by everybody
    if (tid < 16) { (&amax_p_s[ab ^ 1][0][0])[tid] = 0u; (&amax_ds_s[ab ^ 1][0][0])[tid] = 0u; }   // for the next tile
    if (half == 0 && p.rowsum != nullptr) atomicAdd(p.rowsum + qrow, rowsum_ds[0][row] + rowsum_ds[1][row]);
    // ---- pass 2: quantise P and dS with the scale of their block, store both tiles ([q row][128 key bytes], swizzled)
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < CW / 16; ++ch) {
      const int cb = (half * CW + ch * 16) / BK;
      const float amp = __uint_as_float(amax_p_s[ab][rb][cb]), amd = __uint_as_float(amax_ds_s[ab][rb][cb]);
      const float inv_p = amp > 0.f ? __fdividef(127.0f, amp) : 0.f;
      const float inv_ds = amd > 0.f ? __fdividef(127.0f, amd) : 0.f;
      const float cdpi = SAGE ? kPs * inv_ds : magic_scale(sdo_f * sv_c[cb & 3] * kPs * inv_ds);
      const float dpbi = SAGE ? -dlt * kPs * inv_ds : -dlt * kPs * inv_ds - kMagic * cdpi;
      const float2 cdpi2 = make_float2(cdpi, cdpi), ndlti2 = make_float2(dpbi, dpbi);
      const float2 invp2 = make_float2(inv_p, inv_p), magic2 = make_float2(8388608.0f, 8388608.0f);
      uint32_t r2[16];
      tmem_ld16(lane_addr + 128 + half * CW + ch * 16, r2);
      tmem_ld_wait();
      uint32_t wp[4], wd[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t bp[4];
        int bd[4];
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          const int c = q4 * 4 + e;
          const float2 pp = __half22float2(pk[ch * 8 + c / 2]);
          const float2 dq = __fmul2_rn(pp, __ffma2_rn(make_float2(__uint_as_float(r2[c]), __uint_as_float(r2[c + 1])), cdpi2, ndlti2));
          const float2 pq = RN ? __ffma2_rn(pp, invp2, magic2) : __ffma2_rz(pp, invp2, magic2);
          bp[e] = __float_as_uint(pq.x);
          bp[e + 1] = __float_as_uint(pq.y);
          bd[e] = RN ? __float2int_rn(dq.x) : __float2int_rz(dq.x);
          bd[e + 1] = RN ? __float2int_rn(dq.y) : __float2int_rz(dq.y);
        }
        wp[q4] = pack_low_bytes(bp[0], bp[1], bp[2], bp[3]);
        wd[q4] = pack_sat_s8x4(bd[0], bd[1], bd[2], bd[3]);
      }
      const uint32_t off = swz128(row, half * CW + ch * 16);
      sts128(smem_base + L::off_p + off, wp[0], wp[1], wp[2], wp[3]);
      sts128(smem_base + L::off_ds + off, wd[0], wd[1], wd[2], wd[3]);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_bar_sync(2, NT);                                       // P / dS tiles stored; dP read by everybody
    // ---- dV_j += P^T dO, dK_j += dS^T Q: one query-row block at a time (the scales change along the contraction)
    for (int qb = 0; qb < NRB; ++qb) {
      if (leader) {
        tc_fence_after();
        umma_f16_ss(tbase + 256, cdesc_a, cdesc_b, id_cD, 0);
        umma_f16_ss(tbase + 384, cdesc_a, cdesc_b, id_cD, 0);
        for (int k = qb * (BQ / 32); k < (qb + 1) * (BQ / 32); ++k) {
          umma_i8_ss(tbase + 256, umma_smem_desc(a_p + k * 4096, 16, 1024, kSwz128), umma_smem_desc(a_do + k * 32 * D, 16, kSbo, kLay), id_t, 1);
          umma_i8_ss(tbase + 384, umma_smem_desc(a_ds + k * 4096, 16, 1024, kSwz128), umma_smem_desc(a_q + k * 32 * D, 16, kSbo, kLay), id_t, 1);
        }
        umma_commit(&mma_done);
      }
      mbar_wait(&mma_done, n_mma & 1); ++n_mma;
      tc_fence_after();
      // this thread's row is a KEY here: block (qb, kb)
      const float c_dv = __half2float(p.s_do[qblk0 + qb]) * __uint_as_float(amax_p_s[ab][qb][kb]) * (kPs / 127.0f);
      const float c_dk = __uint_as_float(amax_ds_s[ab][qb][kb]) * (1.0f / 127.0f) * __half2float(p.sq[qblk0 + qb]) * p.sm_scale;
#pragma unroll
      for (int ch = 0; ch < DH / 16; ++ch) {
        uint32_t r[16], r2[16];
        tmem_ld16(lane_addr + 256 + half * DH + ch * 16, r);
        tmem_ld16(lane_addr + 384 + half * DH + ch * 16, r2);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          dv_acc[ch * 8 + c] = __ffma2_rn(__fadd2_rn(make_float2(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1])), nM2),
                                          make_float2(c_dv, c_dv), dv_acc[ch * 8 + c]);
          dk_acc[ch * 8 + c] = __ffma2_rn(__fadd2_rn(make_float2(__uint_as_float(r2[2 * c]), __uint_as_float(r2[2 * c + 1])), nM2),
                                          make_float2(c_dk, c_dk), dk_acc[ch * 8 + c]);
        }
      }
      tc_fence_before();
      named_bar_sync(1, NT);                                     // partial columns drained by everybody
    }
    // ---- dQ_t += dS K: one key block at a time; this thread's row is a QUERY row: block (rb, cb)
    float2 dq_acc[DH / 2];
#pragma unroll
    for (int d = 0; d < DH / 2; ++d) dq_acc[d] = make_float2(0.f, 0.f);
    for (int cb = 0; cb < NCB; ++cb) {
      if (leader) {
        tc_fence_after();
        umma_f16_ss(tbase + 256, cdesc_a, cdesc_b, id_cD, 0);
        for (int k = cb * (BK / 32); k < (cb + 1) * (BK / 32); ++k)
          umma_i8_ss(tbase + 256, umma_smem_desc(a_ds + k * 32, 16, 1024, kSwz128), umma_smem_desc(a_k + k * 32 * D, 16, kSbo, kLay), id_q, 1);
        umma_commit(&mma_done);
      }
      mbar_wait(&mma_done, n_mma & 1); ++n_mma;
      tc_fence_after();
      const float c_dq = __uint_as_float(amax_ds_s[ab][rb][cb]) * (1.0f / 127.0f) * __half2float(p.sk[(head_row0 + (size_t)j * 128) / BK + cb]) * p.sm_scale;
#pragma unroll
      for (int ch = 0; ch < DH / 16; ++ch) {
        uint32_t r[16];
        tmem_ld16(lane_addr + 256 + half * DH + ch * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          dq_acc[ch * 8 + c] = __ffma2_rn(__fadd2_rn(make_float2(__uint_as_float(r[2 * c]), __uint_as_float(r[2 * c + 1])), nM2),
                                          make_float2(c_dq, c_dq), dq_acc[ch * 8 + c]);
      }
      tc_fence_before();
      named_bar_sync(2, NT);
    }
    // ---- dQ tile -> fp32 staging (swizzled 32-float atoms) -> TMA reduce-add into the workspace
    if (leader) tma_store_wait_read();                           // the previous tile's reduce-add has read the staging tile
    named_bar_sync(1, NT);
#pragma unroll
    for (int ch = 0; ch < DH / 16; ++ch) {
      const int col = half * DH + ch * 16;
      const uint32_t atom = smem_base + L::off_dq + (col >> 5) * (128 * 128);
#pragma unroll
      for (int c = 0; c < 16; c += 4)
        sts128f(atom + swz128(row, ((col & 31) + c) * 4), dq_acc[ch * 8 + c / 2].x, dq_acc[ch * 8 + c / 2].y, dq_acc[ch * 8 + c / 2 + 1].x,
                dq_acc[ch * 8 + c / 2 + 1].y);
    }
    fence_proxy_async_smem();
    named_bar_sync(2, NT);
    if (leader) {
#pragma unroll
      for (int a = 0; a < D / 32; ++a)
        tma_reduce_add_2d(&tm_dq, smem + L::off_dq + a * (128 * 128), a * 32, (int)head_row0 + t * 128);
      tma_store_commit();
    }
  }
  if (leader) tma_store_wait_all();
  // ---- epilogue: dK_j, dV_j rows (row = key) in fp16
  const size_t krow = head_row0 + (size_t)j * 128 + row;
  __half* dk_dst = p.dk + krow * D + half * DH;
  __half* dv_dst = p.dv + krow * D + half * DH;
#pragma unroll
  for (int d = 0; d < DH; d += 8) {
    uint4 a, b;
    __half2 t2;
    t2 = __float22half2_rn(dk_acc[d / 2]); a.x = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dk_acc[d / 2 + 1]); a.y = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dk_acc[d / 2 + 2]); a.z = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dk_acc[d / 2 + 3]); a.w = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dv_acc[d / 2]); b.x = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dv_acc[d / 2 + 1]); b.y = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dv_acc[d / 2 + 2]); b.z = *reinterpret_cast<uint32_t*>(&t2);
    t2 = __float22half2_rn(dv_acc[d / 2 + 3]); b.w = *reinterpret_cast<uint32_t*>(&t2);
    *reinterpret_cast<uint4*>(dk_dst + d) = a;
    *reinterpret_cast<uint4*>(dv_dst + d) = b;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tbase);
}

// Causal row 0 of every head attends uniformly to all S keys (LEDGER B-1): dV[k] += dO[0] / S for every key k; dQ and
// dK get nothing from it (its P is exactly 0 inside the fused kernel).  dO[0] is taken de-quantised (do_i8 * s_dO).
template <int D>
__global__ void __launch_bounds__(256) int8_bwd_row0_fixup_kernel(const int8_t* __restrict__ do_i8, const __half* __restrict__ s_do,
                                                                  __half* dv, int S) {
  const int bh = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;                 // one 8-element vector of dV[bh]
  if (i >= S * (D / 8)) return;
  const int d0 = (i % (D / 8)) * 8;
  const float c = __half2float(s_do[(size_t)bh * (S / 128)]) / (float)S;
  const int8_t* src = do_i8 + (size_t)bh * S * D + d0;
  uint4* dst = reinterpret_cast<uint4*>(dv + (size_t)bh * S * D) + i;
  uint4 v = *dst;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __half22float2(h[e]);
    h[e] = __floats2half2_rn(fmaf((float)src[2 * e], c, f.x), fmaf((float)src[2 * e + 1], c, f.y));
  }
  *dst = v;
}

template <int D, int NG, bool RN = false, bool CAUSAL = false, bool WS = false>
static int launch_int8_bwd(const void* q_i8, const void* k_i8, const void* v_i8, const void* do_i8, void* dq_ws,
                           const Int8BwdParams& p, int BH, cudaStream_t st) {
  using L = Int8BwdSmem<D>;
  CUtensorMap tq, tk, tv, tdo, tdq;
  const int sw = (D == 128) ? 3 : 2;
  uint64_t dims[2] = {(uint64_t)D, (uint64_t)BH * p.S};
  uint64_t str[1] = {(uint64_t)D};
  uint32_t box[2] = {(uint32_t)D, 128};
  int rc;
  if ((rc = qa_make_tmap(&tq, q_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dims, str, box, sw))) return rc;
  if ((rc = qa_make_tmap(&tk, k_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dims, str, box, sw))) return rc;
  if ((rc = qa_make_tmap(&tv, v_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dims, str, box, sw))) return rc;
  if ((rc = qa_make_tmap(&tdo, do_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dims, str, box, sw))) return rc;
  uint64_t strq[1] = {(uint64_t)D * 4};
  uint32_t boxq[2] = {32, 128};
  if ((rc = qa_make_tmap(&tdq, dq_ws, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, strq, boxq, 3))) return rc;
  dim3 grid((p.S_valid + 127) / 128, BH);                        // k-tiles without a valid key are not launched
  if (WS) {                                                        // warp-specialised: 8 quantise + 8 drain warps
    auto kern = int8_bwd_ws_kernel<D, RN, CAUSAL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
    if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
    kern<<<grid, 512, L::total, st>>>(tq, tk, tv, tdo, tdq, p);
    return qa_check_launch("qa_int8_bwd");
  }
  auto kern = int8_bwd_kernel<D, NG, RN, CAUSAL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  kern<<<grid, 128 * NG, L::total, st>>>(tq, tk, tv, tdo, tdq, p);
  return qa_check_launch("qa_int8_bwd");
}

template <int D, bool RN, bool SAGE = false>
static int launch_int8_bwd_blk(const void* q_i8, const void* k_i8, const void* v_i8, const void* do_i8, void* dq_ws,
                               const Int8BwdParams& p, int BH, int Bq, int Bkv, cudaStream_t st, const void* v_f16 = nullptr,
                               const void* do_f16 = nullptr) {
  using L = std::conditional_t<SAGE, Int8BwdSageSmem<D>, Int8BwdSmem<D>>;
  CUtensorMap tq, tk, tv, tdo, tdq, tdo16;
  const int sw = (D == 128) ? 3 : 2;
  uint64_t dims[2] = {(uint64_t)D, (uint64_t)BH * p.S};
  uint64_t str[1] = {(uint64_t)D};
  uint32_t box[2] = {(uint32_t)D, 128};
  int rc;
  if ((rc = qa_make_tmap(&tq, q_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dims, str, box, sw))) return rc;
  if ((rc = qa_make_tmap(&tk, k_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dims, str, box, sw))) return rc;
  if ((rc = qa_make_tmap(&tdo, do_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dims, str, box, sw))) return rc;
  uint64_t str16[1] = {(uint64_t)D * 2};
  uint32_t box16[2] = {64, 128};
  if (SAGE) {                                                      // fp16 V and dO: [128][64] atoms, 128B swizzle
    if ((rc = qa_make_tmap(&tv, v_f16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dims, str16, box16, 3))) return rc;
    if ((rc = qa_make_tmap(&tdo16, do_f16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dims, str16, box16, 3))) return rc;
  } else {
    if ((rc = qa_make_tmap(&tv, v_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dims, str, box, sw))) return rc;
    tdo16 = tdo;
  }
  uint64_t strq[1] = {(uint64_t)D * 4};
  uint32_t boxq[2] = {32, 128};
  if ((rc = qa_make_tmap(&tdq, dq_ws, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, strq, boxq, 3))) return rc;
  auto kern = int8_bwd_blk_kernel<D, RN, SAGE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  dim3 grid((p.S_valid + 127) / 128, BH);                        // k-tiles without a valid key are not launched
  kern<<<grid, 256, L::total, st>>>(tq, tk, tv, tdo, tdq, tdo16, p, Bq, Bkv);
  return qa_check_launch("qa_int8_bwd");
}

}  // namespace qa

using namespace qa;

// Backward over pre-quantised operands (Bq = Bkv = 128).  dq_ws: fp32 [BH*S, D] zero-initialised accumulation
// workspace (cast to fp16 with qa_cast_f32 afterwards); dk, dv: fp16 [BH*S, D].
#ifdef QA_DEV_TIMELINE
static void* g_int8_bwd_dbg = nullptr;
// Development library only (include/qattn_dev.h, tools/timeline_bwd.py): CTA (0,0) of subsequent qa_int8_bwd launches
// records SM-clock stamps per q-tile into buf ([64 tiles][2][16] int64); NULL switches it off.  Not thread-safe.
extern "C" int qa_debug_set_int8_bwd_timeline(void* buf) {
  g_int8_bwd_dbg = buf;
  return 0;
}
#endif

extern "C" int qa_int8_bwd_ragged(const void* q_i8, const void* k_i8, const void* v_i8, const void* do_i8, const void* sq,
                                  const void* sk, const void* sv, const void* s_do, const void* lse_f32, const void* delta_f32,
                                  void* rowsum_ws_f32, void* dq_ws_f32, void* dk_f16, void* dv_f16, int BH, int S, int S_valid,
                                  int D, int Bq, int Bkv, int flags, void* stream) {
  if (S_valid <= 0 || S_valid > S || S - S_valid >= 128)
    return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: S_valid must be in (S - 128, S] (pad to the next multiple of 128 only)");
  if (S_valid != S && (flags & QA_FLAG_CAUSAL)) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: ragged sequences are built for the non-causal kernels");
  if (flags & ~(QA_FLAG_NEAREST | QA_FLAG_CAUSAL | QA_FLAG_BWD_8WARP)) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: unknown flag bits");
  const int rounding = (flags & QA_FLAG_NEAREST) ? 1 : 0;
  const bool causal = (flags & QA_FLAG_CAUSAL) != 0;
  const bool ws = (flags & QA_FLAG_BWD_8WARP) == 0;     // warp-specialised kernel unless the caller asks for the 8-warp one
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: D must be 64 or 128");
  const bool blk = (Bq != 128 || Bkv != 128);                    // quantisation blocks smaller than the MMA tile
  if ((Bq != 32 && Bq != 64 && Bq != 128) || (Bkv != 32 && Bkv != 64 && Bkv != 128))
    return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: Bq and Bkv must be 32, 64 or 128");
  if (blk && causal) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: causal is built for Bq = Bkv = 128");
  if (BH <= 0 || S <= 0 || S % 128) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: S must be a positive multiple of 128");
  if ((long long)BH * S >= (1ll << 31)) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: BH * S must be below 2^31 (TMA coordinates)");
  if (!q_i8 || !k_i8 || !v_i8 || !do_i8 || !sq || !sk || !sv || !s_do || !lse_f32 || !delta_f32 || !dq_ws_f32 || !dk_f16 || !dv_f16)
    return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: null pointer (only rowsum_ws may be NULL)");
  if (((uintptr_t)q_i8 | (uintptr_t)k_i8 | (uintptr_t)v_i8 | (uintptr_t)do_i8 | (uintptr_t)dq_ws_f32 | (uintptr_t)dk_f16 | (uintptr_t)dv_f16) & 15)
    return qa_fail(QA_ERR_ALIGN, "qa_int8_bwd: 16-byte alignment required");
  Int8BwdParams p;
  p.sq = (const __half*)sq; p.sk = (const __half*)sk; p.sv = (const __half*)sv; p.s_do = (const __half*)s_do;
  p.lse = (const float*)lse_f32; p.delta = (const float*)delta_f32; p.rowsum = (float*)rowsum_ws_f32;
  p.dk = (__half*)dk_f16; p.dv = (__half*)dv_f16;
  p.S = S; p.S_valid = S_valid;
#ifdef QA_DEV_TIMELINE
  p.dbg = (long long*)g_int8_bwd_dbg;
#else
  p.dbg = nullptr;
#endif
  p.sm_scale = (float)(1.0 / sqrt((double)D));
  p.qk_scale = (float)((1.0 / sqrt((double)D)) * 1.44269504);
  cudaStream_t st = (cudaStream_t)stream;
  if (blk) {
    if (rounding) return D == 128 ? launch_int8_bwd_blk<128, true>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, Bq, Bkv, st)
                                  : launch_int8_bwd_blk<64, true>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, Bq, Bkv, st);
    return D == 128 ? launch_int8_bwd_blk<128, false>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, Bq, Bkv, st)
                    : launch_int8_bwd_blk<64, false>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, Bq, Bkv, st);
  }
  if (ws) {
    int rc;
#define QA_WS(DD, RNN, CC) launch_int8_bwd<DD, 2, RNN, CC, true>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, st)
    if (causal && rounding) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: causal is built for truncation mode");
    if (causal) rc = D == 128 ? QA_WS(128, false, true) : QA_WS(64, false, true);
    else if (rounding) rc = D == 128 ? QA_WS(128, true, false) : QA_WS(64, true, false);
    else rc = D == 128 ? QA_WS(128, false, false) : QA_WS(64, false, false);
#undef QA_WS
    if (rc || !causal) return rc;
    dim3 grid((unsigned)((S * (D / 8) + 255) / 256), (unsigned)BH);
    if (D == 128) int8_bwd_row0_fixup_kernel<128><<<grid, 256, 0, st>>>((const int8_t*)do_i8, p.s_do, p.dv, S);
    else int8_bwd_row0_fixup_kernel<64><<<grid, 256, 0, st>>>((const int8_t*)do_i8, p.s_do, p.dv, S);
    return qa_check_launch("qa_int8_bwd(causal row 0)");
  }
  if (causal) {                                                   // instantiated for the default shape only
    if (rounding) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd: causal is built for truncation mode");
    int rc = D == 128 ? launch_int8_bwd<128, 2, false, true>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, st)
                      : launch_int8_bwd<64, 2, false, true>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, st);
    if (rc) return rc;
    dim3 grid((unsigned)((S * (D / 8) + 255) / 256), (unsigned)BH);
    if (D == 128) int8_bwd_row0_fixup_kernel<128><<<grid, 256, 0, st>>>((const int8_t*)do_i8, p.s_do, p.dv, S);
    else int8_bwd_row0_fixup_kernel<64><<<grid, 256, 0, st>>>((const int8_t*)do_i8, p.s_do, p.dv, S);
    return qa_check_launch("qa_int8_bwd(causal row 0)");
  }
  if (rounding == 1)                                              // accuracy mode: instantiated for the default shape only
    return D == 128 ? launch_int8_bwd<128, 2, true>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, st)
                    : launch_int8_bwd<64, 2, true>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, st);
  return D == 128 ? launch_int8_bwd<128, 2>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, st)
                  : launch_int8_bwd<64, 2>(q_i8, k_i8, v_i8, do_i8, dq_ws_f32, p, BH, st);
}

extern "C" int qa_int8_bwd(const void* q_i8, const void* k_i8, const void* v_i8, const void* do_i8, const void* sq,
                           const void* sk, const void* sv, const void* s_do, const void* lse_f32, const void* delta_f32,
                           void* rowsum_ws_f32, void* dq_ws_f32, void* dk_f16, void* dv_f16, int BH, int S, int D,
                           int Bq, int Bkv, int flags, void* stream) {
  return qa_int8_bwd_ragged(q_i8, k_i8, v_i8, do_i8, sq, sk, sv, s_do, lse_f32, delta_f32, rowsum_ws_f32, dq_ws_f32, dk_f16,
                            dv_f16, BH, S, S, D, Bq, Bkv, flags, stream);
}

// SageBwd option: the same backward with dP = dO V^T computed from the UNQUANTISED fp16 dO and V
// (tcgen05 kind::f16, fp32 accumulation) instead of the int8 product the reference uses (attention_int8.py:380-384); every
// other contraction stays int8.  Runs the block kernel (any Bq, Bkv in {32, 64, 128}); non-causal, S a multiple of 128.
extern "C" int qa_int8_bwd_sage(const void* q_i8, const void* k_i8, const void* v_f16, const void* do_i8, const void* do_f16,
                                const void* sq, const void* sk, const void* s_do, const void* lse_f32, const void* delta_f32,
                                void* rowsum_ws_f32, void* dq_ws_f32, void* dk_f16, void* dv_f16, int BH, int S, int D, int Bq,
                                int Bkv, int flags, void* stream) {
  if (flags & ~QA_FLAG_NEAREST) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd_sage: only QA_FLAG_NEAREST is accepted");
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd_sage: D must be 64 or 128");
  if ((Bq != 32 && Bq != 64 && Bq != 128) || (Bkv != 32 && Bkv != 64 && Bkv != 128))
    return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd_sage: Bq and Bkv must be 32, 64 or 128");
  if (BH <= 0 || S <= 0 || S % 128) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd_sage: S must be a positive multiple of 128");
  if ((long long)BH * S >= (1ll << 31)) return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd_sage: BH * S must be below 2^31");
  if (!q_i8 || !k_i8 || !v_f16 || !do_i8 || !do_f16 || !sq || !sk || !s_do || !lse_f32 || !delta_f32 || !dq_ws_f32 || !dk_f16 || !dv_f16)
    return qa_fail(QA_ERR_SHAPE, "qa_int8_bwd_sage: null pointer (only rowsum_ws may be NULL)");
  if (((uintptr_t)q_i8 | (uintptr_t)k_i8 | (uintptr_t)v_f16 | (uintptr_t)do_i8 | (uintptr_t)do_f16 | (uintptr_t)dq_ws_f32 | (uintptr_t)dk_f16 | (uintptr_t)dv_f16) & 15)
    return qa_fail(QA_ERR_ALIGN, "qa_int8_bwd_sage: 16-byte alignment required");
  Int8BwdParams p;
  p.sq = (const __half*)sq; p.sk = (const __half*)sk; p.sv = (const __half*)sk; p.s_do = (const __half*)s_do;   // sv unused
  p.lse = (const float*)lse_f32; p.delta = (const float*)delta_f32; p.rowsum = (float*)rowsum_ws_f32;
  p.dk = (__half*)dk_f16; p.dv = (__half*)dv_f16;
  p.S = S; p.S_valid = S; p.dbg = nullptr;
  p.sm_scale = (float)(1.0 / sqrt((double)D));
  p.qk_scale = (float)((1.0 / sqrt((double)D)) * 1.44269504);
  cudaStream_t st = (cudaStream_t)stream;
  if (flags & QA_FLAG_NEAREST)
    return D == 128 ? launch_int8_bwd_blk<128, true, true>(q_i8, k_i8, nullptr, do_i8, dq_ws_f32, p, BH, Bq, Bkv, st, v_f16, do_f16)
                    : launch_int8_bwd_blk<64, true, true>(q_i8, k_i8, nullptr, do_i8, dq_ws_f32, p, BH, Bq, Bkv, st, v_f16, do_f16);
  return D == 128 ? launch_int8_bwd_blk<128, false, true>(q_i8, k_i8, nullptr, do_i8, dq_ws_f32, p, BH, Bq, Bkv, st, v_f16, do_f16)
                  : launch_int8_bwd_blk<64, false, true>(q_i8, k_i8, nullptr, do_i8, dq_ws_f32, p, BH, Bq, Bkv, st, v_f16, do_f16);
}
