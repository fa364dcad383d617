// Forward-mode JVP attention (reference attention_jvp.py:129-190) for sm_100a.
// Emits O, tO and the log2-LSE.  Six contractions per (query tile, key tile), all tcgen05 kind::f16 with bf16
// operands and fp32 TMEM accumulators (LEDGER J-2; operands are converted fp32 -> bf16 by qa_cast_f32):
//   S  = Q K^T                      (TMEM cols   0..127)
//   tS = tQ K^T + Q tK^T            (TMEM cols 128..255, two chained MMAs into one accumulator)
//   O  += P V                       (TMEM cols 256..256+D, resident accumulator)
//   AB += P tV + H V, H = P*tS      (TMEM cols 256+D.., two chained MMAs: A and B of the reference share one tile)
// Softmax state (m, l, r = rowsum(H)) is fp32 as in the reference.  tO = (AB - r*O) / l,  O = O / l.
// The O and AB accumulators stay RESIDENT in TMEM across k-tiles; the correction warpgroup multiplies them by the
// rescale factor only when a row maximum moved (exp2(0) == 1 exactly otherwise), and the softmax warps keep the S
// tile in registers between the max pass and the exp pass, so steady-state TMEM->register traffic is S + tS only.
#include "qa_ptx.cuh"
#include "qa_host.h"
#include <type_traits>

namespace qa {

constexpr int kJAtom = 128 * 128;

// BN = keys per k-tile: 128 at D = 64; 64 at D = 128 (shared memory: six operand tiles + P + H per stage)
template <int D, int BN>
struct JvpSmem {
  static constexpr int kTileQ = 128 * D * 2;            // Q / tQ tile (128 rows)
  static constexpr int kTile = BN * D * 2;              // K / tK / V / tV tile (BN rows)
  static constexpr int kKAtom = BN * 128;               // one 128-byte column of a BN-row tile
  static constexpr int off_q = 0;                       // Q, tQ
  static constexpr int off_k = off_q + 2 * kTileQ;      // 2 stages x (K, tK)
  static constexpr int off_v = off_k + 4 * kTile;       // 1 stage  x (V, tV)
  static constexpr int total = off_v + 2 * kTile + 1024;
};

struct JvpParams {
  float *O, *tO, *lse;
  int Sq, Sk;
  int Sk_valid;          // keys [Sk_valid, Sk) of every head are zero padding (ragged sequence): weight exactly 0
  float sm_scale, qk_scale;
};

__device__ __forceinline__ uint32_t jpack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int D, int NSPLIT, int BN>
__global__ void __launch_bounds__(NSPLIT == 2 ? 512 : 320, 1)
jvp_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_tq,
               const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_tk,
               const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_tv, JvpParams p) {
  using L = JvpSmem<D, BN>;
  constexpr int NC = BN / NSPLIT;
  static_assert(NC % 32 == 0, "a softmax thread handles a multiple of 32 columns");
  constexpr int kSoftWarps = 4 * NSPLIT;
  constexpr int kProdWarp = kSoftWarps + 4;
  constexpr int kMmaWarp = kSoftWarps + 5;
  // NSPLIT == 2: 16 warps (the last two idle, so that every warpgroup is complete for setmaxnreg): the softmax warps take 168
  // registers and hold S and tS of their columns, which frees the S / tS columns for Q K^T of the next tile right after the loads
  constexpr bool kEarly = (NSPLIT == 2);
  constexpr int kDAtoms = D / 64;
  // TMEM: S [0,BN)  tS [128,128+BN)  O [256,256+D)  AB [256+D,256+2D); P / H (BN/2 columns each) use what is left:
  // D = 64 (BN = 128): [384,448) / [448,512);  D = 128 (BN = 64): [64,96) / [192,224)
  constexpr uint32_t kPCol = (D == 64) ? 384 : 64, kHCol = (D == 64) ? 448 : 192;
  static_assert((D == 64 && BN == 128) || (D == 128 && BN == 64), "TMEM column plan");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t q_full, k_full[2], k_empty[2], v_full, v_empty;
  __shared__ uint64_t s_full, s_empty, p_full, p_empty, o_full, o_ready, sc_full[2], sc_empty[2], fin_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ float row_sc[2][128];
  __shared__ float xmax[2][2][128];
  __shared__ float l_part[2][128], r_part[2][128], m_fin[128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, q0 = blockIdx.x * 128;
  const int nk = (p.Sk_valid + BN - 1) / BN;                  // k-tiles without a valid key are skipped

  if (tid == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    mbar_init(&v_full, 1); mbar_init(&v_empty, 1);
    mbar_init(&s_full, 1); mbar_init(&s_empty, kSoftWarps);
    mbar_init(&p_full, kSoftWarps); mbar_init(&p_empty, 1);
    mbar_init(&o_full, 1); mbar_init(&o_ready, 4);
    for (int b = 0; b < 2; ++b) { mbar_init(&sc_full[b], 4); mbar_init(&sc_empty[b], 4); }
    mbar_init(&fin_full, kSoftWarps);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp < kSoftWarps) {
    if (kEarly) asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");         // 8 x 168 + 8 x 88 registers per lane = 2048
    // =========================== softmax warps ===========================
    const int split = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const int c0 = split * NC;
    float m = -INFINITY, l = 0.f, racc = 0.f;                    // attention_jvp.py:130-134
    for (int j = 0; j < nk; ++j) {
      const uint32_t ph = j & 1;
      mbar_wait(&s_full, ph);
      tc_fence_after();
      float mx = -INFINITY;
      uint32_t sreg[NC];                                          // S row segment stays in registers between the passes
      auto load_s = [&](auto tail) {                               // two instantiations: the key mask costs nothing on full tiles
#pragma unroll
        for (int ch = 0; ch < NC / 32; ++ch) {
          uint32_t r[32];
          tmem_ld32(lane_addr + c0 + ch * 32, r);
          tmem_ld_wait();
          if (decltype(tail)::value) {                            // ragged last k-tile: padded keys get logit -inf, weight 0
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (j * BN + c0 + ch * 32 + i >= p.Sk_valid) r[i] = 0xff800000u;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) { sreg[ch * 32 + i] = r[i]; mx = fmaxf(mx, __uint_as_float(r[i])); }
        }
      };
      if ((j + 1) * BN > p.Sk_valid) load_s(std::true_type{}); else load_s(std::false_type{});
      // NSPLIT == 2: tS of this thread's columns also moves to registers now, so that the S / tS columns are free for Q K^T of the
      // next tile while this tile's exponentials are still being computed (the kernel was a serial chain softmax -> Q K^T ->
      // softmax with every unit below 40 %, profiles/r02_jvp_ncu.txt)
      uint32_t treg[kEarly ? NC : 1];
      if (kEarly) {
#pragma unroll
        for (int ch = 0; ch < NC / 32; ++ch) {
          uint32_t rt[32];
          tmem_ld32(lane_addr + 128 + c0 + ch * 32, rt);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) treg[(kEarly ? ch * 32 : 0) + (kEarly ? i : 0)] = rt[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty);
      }
      if (NSPLIT == 2) {
        xmax[ph][split][row] = mx;
        named_bar_sync(1 + (warp & 3), 32 * NSPLIT);   // only the warps that share these 32 rows meet
        mx = fmaxf(mx, xmax[ph][split ^ 1][row]);
      }
      const float m_new = fmaxf(m, mx * p.qk_scale);             // :155-158
      const float resc = ex2_approx(m - m_new);                   // :164
      m = m_new;
      if (split == 0 && j > 0) {                                  // tile 0 overwrites O / AB: nothing to rescale
        const int sb = (j - 1) & 1;
        mbar_wait(&sc_empty[sb], (((j - 1) >> 1) & 1) ^ 1);
        row_sc[sb][row] = resc;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sc_full[sb]);
      }
      mbar_wait(&p_empty, ph ^ 1);
      float lsum = 0.f, hsum = 0.f;
      // P and H = P * tS (bf16, two per column) go to spare TMEM columns and are consumed from there as the A operands
      // of the three accumulating MMAs (TS mode): no swizzled shared-memory stores, no proxy fence
#pragma unroll
      for (int ch = 0; ch < NC / 32; ++ch) {
        uint32_t rl[32];
        if (!kEarly) {
          tmem_ld32(lane_addr + 128 + c0 + ch * 32, rl);
          tmem_ld_wait();
        }
        const uint32_t* rt = kEarly ? &treg[kEarly ? ch * 32 : 0] : rl;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t wp[8], wh[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int i = g * 16 + e * 2;
            const float p0 = ex2_approx(fmaf(__uint_as_float(sreg[ch * 32 + i]), p.qk_scale, -m_new));        // :160-161
            const float p1 = ex2_approx(fmaf(__uint_as_float(sreg[ch * 32 + i + 1]), p.qk_scale, -m_new));
            const float h0 = p0 * (__uint_as_float(rt[i]) * p.sm_scale);                           // :153, :176
            const float h1 = p1 * (__uint_as_float(rt[i + 1]) * p.sm_scale);
            lsum += p0 + p1;
            hsum += h0 + h1;
            wp[e] = jpack_bf16(p0, p1);
            wh[e] = jpack_bf16(h0, h1);
          }
          const int col = (c0 + ch * 32 + g * 16) / 2;            // two keys per TMEM column
          tmem_st8(lane_addr + kPCol + col, wp);
          tmem_st8(lane_addr + kHCol + col, wh);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      l = l * resc + lsum;                                        // :165
      racc = racc * resc + hsum;                                  // :178
      __syncwarp();
      if (lane == 0) { if (!kEarly) mbar_arrive(&s_empty); mbar_arrive(&p_full); }
    }
    l_part[split][row] = l;
    r_part[split][row] = racc;
    if (split == 0) m_fin[row] = m;
    __syncwarp();
    if (lane == 0) mbar_arrive(&fin_full);
  } else if (warp < kProdWarp) {
    if (kEarly) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    // =========================== correction warpgroup (4 warps, thread = row) ===========================
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    for (int j = 1; j < nk; ++j) {
      const int sb = (j - 1) & 1;
      mbar_wait(&sc_full[sb], ((j - 1) >> 1) & 1);
      const float resc = row_sc[sb][row];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sc_empty[sb]);
      if (__any_sync(0xffffffffu, resc != 1.0f)) {               // O, A, B *= rescale (:167, :173, :180)
        mbar_wait(&o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 2 * D / 32; ++ch) {                // O at 256.., AB at 256 + D..
          uint32_t r[32];
          tmem_ld32(lane_addr + 256 + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * resc);
          tmem_st32(lane_addr + 256 + ch * 32, r);
        }
        tmem_st_wait();
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_ready);
    }
    mbar_wait(&fin_full, 0);
    mbar_wait(&o_full, (nk - 1) & 1);
    tc_fence_after();
    float l = l_part[0][row], rr = r_part[0][row];
    if (NSPLIT == 2) { l += l_part[1][row]; rr += r_part[1][row]; }
    const size_t gr = (size_t)bh * p.Sq + q0 + row;
    const float inv_l = 1.0f / l;
    float* dO_ = p.O + gr * D;
    float* dT_ = p.tO + gr * D;
#pragma unroll
    for (int ch = 0; ch < D / 32; ++ch) {
      uint32_t ro[32], rt[32];
      tmem_ld32(lane_addr + 256 + ch * 32, ro);
      tmem_ld32(lane_addr + 256 + D + ch * 32, rt);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float o[4], t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          o[e] = __uint_as_float(ro[i + e]) * inv_l;                          // :188
          t[e] = (__uint_as_float(rt[i + e]) - rr * o[e]) * inv_l;            // :190  (A + B - r*O) / l
        }
        *reinterpret_cast<float4*>(dO_ + ch * 32 + i) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dT_ + ch * 32 + i) = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    p.lse[gr] = m_fin[row] + log2f(l);                            // :183
  } else if (kEarly && warp > kMmaWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");                      // idle warps of the last warpgroup
  } else if (warp == kProdWarp) {
    if (kEarly) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    // =========================== TMA producer ===========================
    if (elect_one()) {
      mbar_expect_tx(&q_full, 2 * L::kTileQ);
#pragma unroll
      for (int a = 0; a < kDAtoms; ++a) {
        tma_load_2d(smem + L::off_q + a * kJAtom, &tm_q, &q_full, a * 64, bh * p.Sq + q0);
        tma_load_2d(smem + L::off_q + L::kTileQ + a * kJAtom, &tm_tq, &q_full, a * 64, bh * p.Sq + q0);
      }
      for (int j = 0; j < nk; ++j) {
        const int s = j & 1;
        mbar_wait(&k_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&k_full[s], 2 * L::kTile);
#pragma unroll
        for (int a = 0; a < kDAtoms; ++a) {
          tma_load_2d(smem + L::off_k + s * 2 * L::kTile + a * L::kKAtom, &tm_k, &k_full[s], a * 64, bh * p.Sk + j * BN);
          tma_load_2d(smem + L::off_k + s * 2 * L::kTile + L::kTile + a * L::kKAtom, &tm_tk, &k_full[s], a * 64, bh * p.Sk + j * BN);
        }
        mbar_wait(&v_empty, (j & 1) ^ 1);
        mbar_expect_tx(&v_full, 2 * L::kTile);
#pragma unroll
        for (int a = 0; a < kDAtoms; ++a) {
          tma_load_2d(smem + L::off_v + a * L::kKAtom, &tm_v, &v_full, a * 64, bh * p.Sk + j * BN);
          tma_load_2d(smem + L::off_v + L::kTile + a * L::kKAtom, &tm_tv, &v_full, a * 64, bh * p.Sk + j * BN);
        }
      }
    }
  } else {
    if (kEarly) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc(1, 1, 1, 0, 0, 128, BN);         // bf16 x bf16 -> f32, K-major
      constexpr uint32_t idesc_pv = umma_idesc(1, 1, 1, 0, 1, 128, D);          // B MN-major
      constexpr uint32_t idesc_pv2 = umma_idesc(1, 1, 1, 0, 1, 128, 2 * D);     // B = [V | tV]: 2 D columns
      static_assert(L::kTile == kDAtoms * L::kKAtom, "tV must continue V's N atoms");
      const uint32_t q_addr = smem_u32(smem + L::off_q), tq_addr = q_addr + L::kTileQ;
      const uint32_t v_addr = smem_u32(smem + L::off_v);           // tV follows at v_addr + L::kTile
      auto issue_pv = [&](int t) {
        const uint32_t ph = t & 1;
        mbar_wait(&v_full, ph);
        if (t > 0) mbar_wait(&o_ready, (t - 1) & 1);              // accumulators rescaled (if needed) for tile t
        mbar_wait(&p_full, ph);
        tc_fence_after();
        // A TS-mode MMA costs ~153 clk whatever its N (profiles/r02_mma_rate.txt), and this kernel is bound by them (D = 64: 24 of
        // them per tile).  V and tV are adjacent in shared memory with the same layout (tV = V + one tile = the next N atoms of an
        // MN-major B operand) and O / AB are adjacent in TMEM, so [O | AB] += P [V | tV] is ONE instruction with N = 2 D.
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          const uint64_t vd = umma_smem_desc(v_addr + k * 2048, L::kKAtom, 1024, kSwz128);
          umma_f16_ts(tbase + 256, tbase + kPCol + k * 8, vd, idesc_pv2, (t > 0) || (k > 0));      // [O | AB] += P [V | tV]
          umma_f16_ts(tbase + 256 + D, tbase + kHCol + k * 8, vd, idesc_pv, 1);                    //  AB      += H V
        }
        umma_commit(&o_full);
        umma_commit(&v_empty);
        umma_commit(&p_empty);
      };
      mbar_wait(&q_full, 0);
      for (int j = 0; j < nk; ++j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        mbar_wait(&s_empty, (j & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + L::off_k + s * 2 * L::kTile), tk_addr = k_addr + L::kTile;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t o = (k >> 2) * kJAtom + (k & 3) * 32;
          const uint32_t ok = (k >> 2) * L::kKAtom + (k & 3) * 32;
          const uint64_t qd = umma_smem_desc(q_addr + o, 16, 1024, kSwz128), tqd = umma_smem_desc(tq_addr + o, 16, 1024, kSwz128);
          const uint64_t kd = umma_smem_desc(k_addr + ok, 16, 1024, kSwz128), tkd = umma_smem_desc(tk_addr + ok, 16, 1024, kSwz128);
          umma_f16_ss(tbase + 0, qd, kd, idesc_qk, k > 0);          // S
          umma_f16_ss(tbase + 128, tqd, kd, idesc_qk, k > 0);       // tS  = tQ K^T
          umma_f16_ss(tbase + 128, qd, tkd, idesc_qk, 1);           //     + Q tK^T
        }
        umma_commit(&s_full);
        umma_commit(&k_empty[s]);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(nk - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc<512>(tbase);
}

template <int D, int NSPLIT, int BN>
static int launch_jvp(const void* const* in6, const JvpParams& p, int BH, cudaStream_t st) {
  using L = JvpSmem<D, BN>;
  CUtensorMap tm[6];
  uint64_t str[1] = {(uint64_t)D * 2};
  for (int i = 0; i < 6; ++i) {
    uint64_t dims[2] = {(uint64_t)D, (uint64_t)BH * (i < 2 ? p.Sq : p.Sk)};
    uint32_t box[2] = {64, (uint32_t)(i < 2 ? 128 : BN)};
    int rc = qa_make_tmap(&tm[i], in6[i], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3);
    if (rc) return rc;
  }
  auto kern = jvp_fwd_kernel<D, NSPLIT, BN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  dim3 grid(p.Sq / 128, BH);
  kern<<<grid, NSPLIT == 2 ? 512 : 320, L::total, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
  return qa_check_launch("qa_jvp_fwd");
}

}  // namespace qa

using namespace qa;

// q, tq: bf16 [BH*Sq, D]; k, tk, v, tv: bf16 [BH*Sk, D]; O, tO: fp32 [BH*Sq, D]; lse: fp32 [BH*Sq].
// Ragged sequences (hl.tile clamps the last tile, attention_jvp.py:120,137): buffers zero-padded per head to Sq / Sk
// (multiples of 128), keys [Sk_valid, Sk) have weight exactly 0.
extern "C" int qa_jvp_fwd_ragged(const void* q, const void* tq, const void* k, const void* tk, const void* v, const void* tv,
                                 void* O_f32, void* tO_f32, void* lse_f32, int BH, int Sq, int Sk, int Sk_valid, int D, int nsplit,
                                 void* stream) {
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_jvp_fwd: D must be 64 or 128");
  if (Sq % 128 || Sk % 128 || Sq <= 0 || Sk <= 0 || BH <= 0) return qa_fail(QA_ERR_SHAPE, "qa_jvp_fwd: Sq, Sk must be positive multiples of 128");
  if (Sk_valid <= Sk - 128 || Sk_valid > Sk) return qa_fail(QA_ERR_SHAPE, "qa_jvp_fwd: Sk_valid must lie in (Sk - 128, Sk]");
  const void* in6[6] = {q, tq, k, tk, v, tv};
  for (int i = 0; i < 6; ++i)
    if (!in6[i] || ((uintptr_t)in6[i] & 15)) return qa_fail(QA_ERR_ALIGN, "qa_jvp_fwd: null or not 16-byte aligned pointer");
  if (!O_f32 || !tO_f32 || !lse_f32) return qa_fail(QA_ERR_ALIGN, "qa_jvp_fwd: null output pointer");
  JvpParams p;
  p.O = (float*)O_f32; p.tO = (float*)tO_f32; p.lse = (float*)lse_f32; p.Sq = Sq; p.Sk = Sk; p.Sk_valid = Sk_valid;
  p.sm_scale = (float)(1.0 / sqrt((double)D));
  p.qk_scale = (float)((1.0 / sqrt((double)D)) * 1.44269504);
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 128) return nsplit == 1 ? launch_jvp<128, 1, 64>(in6, p, BH, st) : launch_jvp<128, 2, 64>(in6, p, BH, st);
  return nsplit == 2 ? launch_jvp<64, 2, 128>(in6, p, BH, st) : launch_jvp<64, 1, 128>(in6, p, BH, st);
}

extern "C" int qa_jvp_fwd(const void* q, const void* tq, const void* k, const void* tk, const void* v, const void* tv,
                          void* O_f32, void* tO_f32, void* lse_f32, int BH, int Sq, int Sk, int D, int nsplit, void* stream) {
  return qa_jvp_fwd_ragged(q, tq, k, tk, v, tv, O_f32, tO_f32, lse_f32, BH, Sq, Sk, Sk, D, nsplit, stream);
}
