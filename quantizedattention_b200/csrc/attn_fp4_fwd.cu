// NVFP4 (microscaling) attention forward for sm_100a (the SageAttention3 headline feature the reference
// names but does not ship, /root/reference/README.md:48-54).  Contract: oracle/fp4_ref.py.
//
// Both contractions run on tcgen05.mma kind::mxf4nvf4.block_scale.block16 (e2m1 operands, one ue4m3 scale per 16 elements
// along the contraction axis, fp32 accumulation in TMEM); the operand, scale-factor and instruction-descriptor conventions
// were pinned on the hardware by qa_probe_mma_bs (tests/test_probe_gpu.py, profiles/r02_fp4_probe.txt).
//   S = Q4 K4^T      A = Q tile, B = K tile: rows of D/2 = 64 bytes, 64-byte swizzle, two K = 64 steps; scale factors of Q / K
//                    arrive as 512-byte atoms by TMA and are copied shared memory -> TMEM by tcgen05.cp.32x128b.warpx4
//   O += P4 V4       A = P from TMEM (TS mode: 8 e2m1 per 32-bit column, written by the exp warps into P's own columns),
//                    B = V^T tile [D rows, 128 keys] (V is stored transposed: 4-bit operands are K-major only);
//                    P's scale factors go through a shared-memory atom and tcgen05.cp (they must be replicated over
//                    the four TMEM lane quadrants, which a thread cannot write)
// Because the microscales travel with the operands, the fp32 accumulator spans k-tiles (unlike the int8 / fp8 path, whose
// per-tile P and V scales force a drain every k-tile): O stays resident in TMEM and is rescaled only when a row maximum moves.
// One CTA = one 128-row query tile of one head.  Options: strict causal mask (tile skipping, diagonal masking, row-0 fix-up),
// ragged key lengths (Sk_valid), sm_scale.
// Default kernel (variant 0), 20 warps: 0-7 exp (two per 32-row group, alternating tiles: the XU-bound stage), 8-11 running maximum
// (a tile ahead) and epilogue, 12-15 correction (rescale of O), 16 TMA producer, 17 / 18 MMA issuers (P V / Q K^T), 19 idle.
// TMEM (512 columns): S[2] at 0 / 128, O at 256, scale factors from 384: Q 8, K 2 x 8, V 2 x 8, P 2 x 8 columns, P[2] at 440 / 456.
#include "qa_ptx.cuh"
#include "qa_host.h"
#include <cuda_fp4.h>
#include <cuda_fp8.h>

namespace qa {

constexpr int kFp4D = 128;

template <int STAGES>
struct Fp4FwdSmem {
  static constexpr int kTile = 128 * 64;                       // 128 rows x 128 e2m1 = 8 KB
  static constexpr int kSf = 1024;                             // two 512-byte scale-factor atoms (K steps of 64)
  static constexpr int off_q = 0;
  static constexpr int off_k = off_q + kTile;
  static constexpr int off_v = off_k + STAGES * kTile;
  static constexpr int off_sfq = off_v + STAGES * kTile;
  static constexpr int off_sfk = off_sfq + kSf;
  static constexpr int off_sfv = off_sfk + STAGES * kSf;
  static constexpr int off_sfp = off_sfv + STAGES * kSf;       // [2] written by the softmax warps
  static constexpr int total = off_sfp + 2 * kSf + 1024;       // + alignment slack
};

struct Fp4FwdParams {
  const float *sgq, *sgk, *sgv;      // [BH] second-level scales
  __half* O;                         // [BH*Sq, 128] fp16
  float* lse;                        // [BH*Sq] log2 domain
  int Sq, Sk, BH;
  int Sk_valid;                      // keys [Sk_valid, Sk) of every head are zero padding (ragged sequence): weight exactly 0
  float qk_scale;                    // sm_scale * log2(e)
};

__device__ __forceinline__ void tmem_cp_sf(uint32_t taddr, uint32_t saddr) {   // one 512-byte atom -> 4 TMEM columns, all lane quadrants
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(umma_smem_desc(saddr, 0, 128, kSwzNone)) : "memory");
}
__device__ __forceinline__ float fp4_e4m3_to_float(uint32_t c) {
  const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)c, __NV_E4M3);
  return __half2float(*reinterpret_cast<const __half*>(&h));
}
// eight floats -> eight e2m1 codes in one word (element 0 in the low nibble), round to nearest even, saturating
__device__ __forceinline__ uint32_t fp4_pack8(const float2 (&y)[4]) {
  uint32_t w;
  asm("{\n\t.reg .b8 t0, t1, t2, t3;\n\t"
      "cvt.rn.satfinite.e2m1x2.f32 t0, %2, %1;\n\t"
      "cvt.rn.satfinite.e2m1x2.f32 t1, %4, %3;\n\t"
      "cvt.rn.satfinite.e2m1x2.f32 t2, %6, %5;\n\t"
      "cvt.rn.satfinite.e2m1x2.f32 t3, %8, %7;\n\t"
      "mov.b32 %0, {t0, t1, t2, t3};\n\t}"
      : "=r"(w)
      : "f"(y[0].x), "f"(y[0].y), "f"(y[1].x), "f"(y[1].y), "f"(y[2].x), "f"(y[2].y), "f"(y[3].x), "f"(y[3].y));
  return w;
}

template <int STAGES, bool CAUSAL>
__global__ void __launch_bounds__(640, 1)
fp4_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
               const __grid_constant__ CUtensorMap tm_vt, const __grid_constant__ CUtensorMap tm_sfq,
               const __grid_constant__ CUtensorMap tm_sfk, const __grid_constant__ CUtensorMap tm_sfv, Fp4FwdParams p) {
  using L = Fp4FwdSmem<STAGES>;
  constexpr int D = kFp4D;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t q_full, kv_full[STAGES], kv_empty[STAGES], s_full[2], s_free[2], p_full[2], o_full[2], o_ready[2], fin_full;
  __shared__ uint64_t mx_full[8][4];          // [tile & 7][row group]: running maximum and rescale factor of the tile are published
  __shared__ float2 prm[8][128];              // (m', 2^(m - m')) per row; a slot is rewritten eight tiles later
  __shared__ float l_part[2][128];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // CAUSAL: the strict mask of the reference's baseline (key < query, attention_int8.py:465-473), as on the int8 path: a query
  // tile meets the k-tiles up to its own (heaviest tiles first); row 0 of a head sees no key and is written by fp4_row0_fixup_kernel
  // (causal: 1-D grid in groups of 16 heads, inside a group the heaviest tiles of all its heads first - global heaviest-first
  //  would spread the K / V working set of the running CTAs over every head and out of the L2)
  int rank_ = 0, head_ = 0;
  if (CAUSAL) qa_group_order((int)blockIdx.x, p.BH, p.Sq / 128, 16, rank_, head_);
  const int bh = CAUSAL ? head_ : (int)blockIdx.y, q0 = (CAUSAL ? p.Sq / 128 - 1 - rank_ : (int)blockIdx.x) * 128;
  const int nk_valid = (p.Sk_valid + 127) / 128;                  // ragged: k-tiles without a valid key are skipped
  const int nk = CAUSAL ? min(nk_valid, q0 / 128 + 1) : nk_valid;
  const int ktail = p.Sk_valid - (nk_valid - 1) * 128;           // valid keys of the last tile (128 unless the sequence is ragged)
  const int jd = q0 / 128;                                       // CAUSAL: the diagonal k-tile (local column < local row is visible)

  if (tid == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 2); }   // K side + V side
    for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_free[b], 4); mbar_init(&p_full[b], 4); mbar_init(&o_full[b], 1); mbar_init(&o_ready[b], 4); }
    for (int r = 0; r < 8; ++r)
      for (int qd = 0; qd < 4; ++qd) mbar_init(&mx_full[r][qd], 1);
    mbar_init(&fin_full, 8);
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  constexpr uint32_t kSfQ = 384, kSfK = 392, kSfV = 408, kSfP = 424;     // TMEM columns of the scale factors
  constexpr uint32_t kP = 440;                                   // P[2]: 16 columns each (128 e2m1 per row), separate from S so that an S buffer
                                                                 // is refilled as soon as the exp warps hold its logits in registers
  const int qd = warp & 3;                                        // TMEM lane quadrant of this warp
  const int row = qd * 32 + lane;
  const uint32_t lane_addr = tbase + ((uint32_t)(qd * 32) << 16);

  if (warp < 8) {
    // =========================== exp warps: exp2 + microscaling, the XU-bound stage ===========================
    // The two exp warps of a 32-row group are DE-PHASED: warp e handles the tiles j = e (mod 2), all 128 keys, so that one of
    // them is in its XU-bound inner loop (MUFU.EX2, F2FP) while the other waits, loads or hands P over.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");           // 8 x 136 + 4 x 88 + 4 x 80 + 4 x 40 = 1920 = 20 x 96
    const int e = warp >> 2;
    const float c = p.sgq[bh] * p.sgk[bh] * p.qk_scale;
    const float2 c2 = make_float2(c, c);
    float l = 0.f;
    for (int j = e; j < nk; j += 2) {
      const int sb = j & 1;
      mbar_wait(&mx_full[j & 7][qd], (j >> 3) & 1);
      tc_fence_after();
      const float2 pr = prm[j & 7][row];
      float resc = pr.y;
      if (j >= 2) resc *= prm[(j - 1) & 7][row].y;                 // the tile the other exp warp of this row group handled
      const float2 nm2 = make_float2(-pr.x, -pr.x);
      float2 ls2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {                             // 64 keys (one K step of P V) at a time
        uint32_t r[64];
        tmem_ld64(lane_addr + sb * 128 + hf * 64, r);
        tmem_ld_wait();
        if (hf == 1) {                                             // both halves of the logits are in registers: S[sb] may be refilled
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[sb]);
        }
        uint32_t pw[8], sfw = 0u;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          float2 pe[8];
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(r[b * 16 + 2 * i]), __uint_as_float(r[b * 16 + 2 * i + 1])), c2, nm2);
            pe[i] = make_float2(ex2_approx(x.x), ex2_approx(x.y));
            if (CAUSAL && j == jd) {                               // diagonal tile: keep key < query
              const int col = hf * 64 + b * 16 + 2 * i;
              if (col >= row) pe[i].x = 0.f;
              if (col + 1 >= row) pe[i].y = 0.f;
            }
            if (ktail < 128 && j == nk_valid - 1) {                // ragged last tile: padding keys have weight 0
              const int col = hf * 64 + b * 16 + 2 * i;
              if (col >= ktail) pe[i].x = 0.f;
              if (col + 1 >= ktail) pe[i].y = 0.f;
            }
            if (i & 1) a1 = fmaxf(a1, fmaxf(pe[i].x, pe[i].y)); else a0 = fmaxf(a0, fmaxf(pe[i].x, pe[i].y));
            ls2 = __fadd2_rn(ls2, pe[i]);
          }
          const uint32_t sc = (uint32_t)__nv_cvt_float_to_fp8(fmaxf(a0, a1) * 448.0f, __NV_SATFINITE, __NV_E4M3);
          const float sf = fp4_e4m3_to_float(sc);
          const float inv = sf > 0.f ? __fdividef(2688.0f, sf) : 0.f;
          const float2 inv2 = make_float2(inv, inv);
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            float2 y[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) y[i] = __fmul2_rn(pe[h8 * 4 + i], inv2);
            pw[2 * b + h8] = fp4_pack8(y);
          }
          sfw |= sc << (8 * b);
        }
        if (hf == 0 && j >= 2) {                                   // P[sb] and its scale-factor atom were last read by P V of tile j-2
          mbar_wait(&o_full[sb], ((j - 2) >> 1) & 1);
          tc_fence_after();
        }
        tmem_st8(lane_addr + kP + sb * 16 + hf * 8, pw);
        *reinterpret_cast<uint32_t*>(smem + L::off_sfp + sb * L::kSf + hf * 512 + 16 * lane + 4 * qd) = sfw;
      }
      l = l * resc + (ls2.x + ls2.y);
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
    }
    if (nk > 0 && ((nk - 1) & 1) != e) {                           // the last tile belonged to the other warp: its rescale
      mbar_wait(&mx_full[(nk - 1) & 7][qd], ((nk - 1) >> 3) & 1);
      l *= prm[(nk - 1) & 7][row].y;
    }
    l_part[e][row] = l;
    __syncwarp();
    if (lane == 0) mbar_arrive(&fin_full);
  } else if (warp < 12) {
    // =========================== maximum warps (thread = row): running maximum of every tile, ahead of the exp warps ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    const float c = p.sgq[bh] * p.sgk[bh] * p.qk_scale;           // >= 0: the maximum commutes with the scaling
    float m = -INFINITY;
    for (int j = 0; j < nk; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t r[64];
        tmem_ld64(lane_addr + sb * 128 + ch * 64, r);
        tmem_ld_wait();
        if (CAUSAL && j == jd) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (ch * 64 + i >= row) r[i] = 0xff800000u;             // masked logit: -inf
        }
        if (ktail < 128 && j == nk_valid - 1) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (ch * 64 + i >= ktail) r[i] = 0xff800000u;           // padding key
        }
#pragma unroll
        for (int i = 0; i < 64; i += 8)
#pragma unroll
          for (int a = 0; a < 4; ++a) mx4[a] = fmaxf(mx4[a], fmaxf(__uint_as_float(r[i + 2 * a]), __uint_as_float(r[i + 2 * a + 1])));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m, (CAUSAL && mx == -INFINITY) ? -INFINITY : mx * c);   // (a row without a visible key: -inf * 0 would be NaN)
      // rescale factor: 0 on the first tile (m = -inf); 1 while a causal row has not seen any key yet ((-inf) - (-inf))
      prm[j & 7][row] = make_float2(m_new, (CAUSAL && m_new == -INFINITY) ? 1.0f : ex2_approx(m - m_new));
      m = m_new;
      tc_fence_before();                                           // the exp warps overwrite these columns with P
      __syncwarp();
      if (lane == 0) mbar_arrive(&mx_full[j & 7][qd]);
    }
    // ---- epilogue: O * sgv / (2688 * l), log2-LSE
    mbar_wait(&fin_full, 0);
    const float l = l_part[0][row] + l_part[1][row];
    mbar_wait(&o_full[(nk - 1) & 1], ((nk - 1) >> 1) & 1);
    tc_fence_after();
    const size_t gr = (size_t)bh * p.Sq + q0 + row;
    const float sc_o = __fdividef(p.sgv[bh], 2688.0f * l);
    __half* dst = p.O + gr * D;
    const bool row0 = CAUSAL && q0 + row == 0;                     // no visible key (l = 0): written by fp4_row0_fixup_kernel
#pragma unroll
    for (int ch = 0; ch < D / 32; ++ch) {
      uint32_t o[32];
      tmem_ld32(lane_addr + 256 + ch * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 v;
        __half2 t;
        t = __floats2half2_rn(__uint_as_float(o[i]) * sc_o, __uint_as_float(o[i + 1]) * sc_o); v.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(__uint_as_float(o[i + 2]) * sc_o, __uint_as_float(o[i + 3]) * sc_o); v.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(__uint_as_float(o[i + 4]) * sc_o, __uint_as_float(o[i + 5]) * sc_o); v.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2half2_rn(__uint_as_float(o[i + 6]) * sc_o, __uint_as_float(o[i + 7]) * sc_o); v.w = *reinterpret_cast<uint32_t*>(&t);
        if (!row0) *reinterpret_cast<uint4*>(dst + ch * 32 + i) = v;
      }
    }
    if (!row0) p.lse[gr] = m + log2f(l);
  } else if (warp < 16) {
    // =========================== correction warps (thread = row): O *= 2^(m - m') between P V(j-1) and P V(j) ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    for (int j = 1; j < nk; ++j) {
      mbar_wait(&mx_full[j & 7][qd], (j >> 3) & 1);
      const float resc = prm[j & 7][row].y;
      if (__any_sync(0xffffffffu, resc != 1.0f)) {                 // only when a row maximum of the warp's 32 rows moved
        mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);       // P V of tile j-1 has landed in TMEM
        tc_fence_after();
        const float2 rs2 = make_float2(resc, resc);
#pragma unroll
        for (int ch = 0; ch < D / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(lane_addr + 256 + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 t = __fmul2_rn(make_float2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), rs2);
            o[i] = __float_as_uint(t.x); o[i + 1] = __float_as_uint(t.y);
          }
          tmem_st32(lane_addr + 256 + ch * 32, o);
        }
        tmem_st_wait();
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_ready[j & 1]);
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");            // warps 16-19 (19 idle): one instruction for the warpgroup
    if (warp == 16) {
      // =========================== TMA producer ===========================
      if (elect_one()) {
        tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_vt);
        mbar_expect_tx(&q_full, L::kTile + L::kSf);
        tma_load_2d(smem + L::off_q, &tm_q, &q_full, 0, bh * p.Sq + q0);
        tma_load_2d(smem + L::off_sfq, &tm_sfq, &q_full, 0, (bh * p.Sq + q0) / 128);
        for (int j = 0; j < nk; ++j) {
          const int s = j % STAGES;
          mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
          mbar_expect_tx(&kv_full[s], 2 * L::kTile + 2 * L::kSf);
          tma_load_2d(smem + L::off_k + s * L::kTile, &tm_k, &kv_full[s], 0, bh * p.Sk + j * 128);
          tma_load_2d(smem + L::off_sfk + s * L::kSf, &tm_sfk, &kv_full[s], 0, bh * (p.Sk / 128) + j);
          tma_load_2d(smem + L::off_v + s * L::kTile, &tm_vt, &kv_full[s], j * 64, bh * D);
          tma_load_2d(smem + L::off_sfv + s * L::kSf, &tm_sfv, &kv_full[s], 0, bh * (p.Sk / 128) + j);
        }
      }
    } else if (warp == 18) {
      // =========================== MMA issuer 1: S = Q K^T ===========================
      // Two issuing threads: the logits run ahead as far as the S buffers allow (gated only by s_free), independent of the
      // thread below, which spends most of its time waiting for P.  A tcgen05.cp costs the issuing thread ~100 clk
      // (tools/ubench/cp_rate.cu), as much as an MMA.
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc_bs(1, 1, 0, 0, 128, 128, 0);          // e2m1 x e2m1, K-major, ue4m3 scales, N = 128 keys
        const uint32_t q_addr = smem_u32(smem + L::off_q);
        mbar_wait(&q_full, 0);
        tc_fence_after();
        tmem_cp_sf(tbase + kSfQ, smem_u32(smem + L::off_sfq));
        tmem_cp_sf(tbase + kSfQ + 4, smem_u32(smem + L::off_sfq) + 512);
        for (int j = 0; j < nk; ++j) {
          const int s = j % STAGES, sb = j & 1;
          if (j >= 2) mbar_wait(&s_free[sb], ((j - 2) >> 1) & 1);   // the exp warps hold the logits of tile j-2 (so S(j-2) is complete too)
          mbar_wait(&kv_full[s], (j / STAGES) & 1);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(smem + L::off_k + s * L::kTile), sfk = smem_u32(smem + L::off_sfk + s * L::kSf);
          tmem_cp_sf(tbase + kSfK + sb * 8, sfk);
          tmem_cp_sf(tbase + kSfK + sb * 8 + 4, sfk + 512);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_nvf4_ss(tbase + sb * 128, umma_smem_desc(q_addr + k * 32, 16, 512, kSwz64), umma_smem_desc(k_addr + k * 32, 16, 512, kSwz64),
                         idesc, tbase + kSfQ + k * 4, tbase + kSfK + sb * 8 + k * 4, k > 0);
          umma_commit(&s_full[sb]);
          umma_commit(&kv_empty[s]);                // K side of the stage (the V side: issuer 2)
        }
      }
    } else if (warp == 17) {
      // =========================== MMA issuer 2: O += P V ===========================
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc_bs(1, 1, 0, 0, 128, 128, 0);          // N = D
        auto stage_sfv = [&](int j) {                                                // scale factors of V_j -> TMEM, one tile ahead
          mbar_wait(&kv_full[j % STAGES], (j / STAGES) & 1);
          tc_fence_after();
          const uint32_t sfv = smem_u32(smem + L::off_sfv + (j % STAGES) * L::kSf);
          tmem_cp_sf(tbase + kSfV + (j & 1) * 8, sfv);
          tmem_cp_sf(tbase + kSfV + (j & 1) * 8 + 4, sfv + 512);
        };
        stage_sfv(0);
        for (int j = 0; j < nk; ++j) {
          const int s = j % STAGES, sb = j & 1;
          if (j + 1 < nk) stage_sfv(j + 1);         // buffer (j+1) & 1: P V(j-1) is ahead of it in this thread's order
          mbar_wait(&p_full[sb], (j >> 1) & 1);     // P and its scale factors are in place
          if (j > 0) mbar_wait(&o_ready[sb], ((j - 1) >> 1) & 1);   // ... and O carries the rescale of this tile
          tc_fence_after();
          const uint32_t v_addr = smem_u32(smem + L::off_v + s * L::kTile);
          const uint32_t sfp = smem_u32(smem + L::off_sfp + sb * L::kSf);
          tmem_cp_sf(tbase + kSfP + sb * 8, sfp);
          tmem_cp_sf(tbase + kSfP + sb * 8 + 4, sfp + 512);
#pragma unroll
          for (int k = 0; k < 2; ++k)                                               // O += P_j V_j, keys 64k .. 64k + 63
            umma_nvf4_ts(tbase + 256, tbase + kP + sb * 16 + k * 8, umma_smem_desc(v_addr + k * 32, 16, 512, kSwz64), idesc,
                         tbase + kSfP + sb * 8 + k * 4, tbase + kSfV + sb * 8 + k * 4, (j > 0) || (k > 0));
          umma_commit(&o_full[sb]);
          umma_commit(&kv_empty[s]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc<512>(tbase);
}

// Causal row 0 of every head sees no key: the reference's baseline (finite fill value) makes it the uniform average over ALL keys
// of V, here of the de-quantised V (as int8_row0_fixup_kernel); lse = -128 + log2(S).  grid = BH, thread = output column d.
__global__ void __launch_bounds__(128) fp4_row0_fixup_kernel(const uint8_t* __restrict__ vt4, const uint8_t* __restrict__ sfv,
                                                             const float* __restrict__ sgv, __half* O, float* lse, int S, int S_valid) {
  __shared__ float lut[16];
  if (threadIdx.x < 16) {
    const float mag[8] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f};
    lut[threadIdx.x] = (threadIdx.x & 8) ? -mag[threadIdx.x & 7] : mag[threadIdx.x & 7];
  }
  __syncthreads();
  const int bh = blockIdx.x, d = threadIdx.x;
  const uint8_t* rowp = vt4 + ((size_t)bh * kFp4D + d) * (S / 2);
  float acc = 0.f;
  for (int j = 0; j < S / 128; ++j) {
    const uint8_t* sf = sfv + ((size_t)bh * (S / 128) + j) * 1024 + 16 * (d % 32) + 4 * (d / 32);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const uint2 w = *reinterpret_cast<const uint2*>(rowp + (size_t)j * 64 + b * 8);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sum += lut[(w.x >> (4 * i)) & 15];
        sum += lut[(w.y >> (4 * i)) & 15];
      }
      acc = fmaf(sum, fp4_e4m3_to_float(sf[(b / 4) * 512 + (b % 4)]), acc);
    }
  }
  O[(size_t)bh * S * kFp4D + d] = __float2half_rn(acc * sgv[bh] / (float)S_valid);     // padded V rows are zero
  if (d == 0) lse[(size_t)bh * S] = -128.0f + log2f((float)S_valid);
}

// ---------------------------------------------------------------------------------------------------------
// Variant 1: TWO CTAs per SM, a simpler schedule kept for comparison (1010 - 1030 TFLOPS vs 1060 of the kernel above): four
// softmax warps (thread = row) do maximum, exp2 and microscaling in one pass; two independent CTAs on an SM interleave their
// latency chains.  To fit, a CTA owns 256 TMEM columns and 64 KB of shared memory and advances in steps of 64 keys:
//   TMEM: S [0,64)  P[2] [64,80)  scale factors Q [80,88) K[2] [88,104) V[2] [104,120) P[2] [120,128)  O [128,256)
//   S = Q K_h^T is an N = 64 MMA; the scales of keys 64h.. are columns 2h, 2h + 1 of the tile's 128-row atom (probe case
//   nvf4_blockscaled_n64_upper_half); P V is one K = 64 MMA per step.  S is single-buffered: the issuer refills it as soon as
//   the softmax warps hold the logits in registers (s_free); P and its scale factors are double-buffered.
// 12 warps: 0-3 softmax (thread = row, setmaxnreg 128), 4-7 correction, 8 TMA producer, 9 MMA issuer, 10-11 idle.
// ---------------------------------------------------------------------------------------------------------
template <int STAGES>
struct Fp4Fwd2Smem {
  static constexpr int kTile = 128 * 64;
  static constexpr int kSf = 1024;
  static constexpr int off_q = 0;
  static constexpr int off_k = off_q + kTile;
  static constexpr int off_v = off_k + STAGES * kTile;
  static constexpr int off_sfq = off_v + STAGES * kTile;
  static constexpr int off_sfk = off_sfq + kSf;
  static constexpr int off_sfv = off_sfk + STAGES * kSf;
  static constexpr int off_sfp = off_sfv + STAGES * kSf;       // [2] x 512 B
  static constexpr int total = off_sfp + 1024 + 1024;
};

template <int STAGES>
__global__ void __launch_bounds__(384, 2)
fp4_fwd2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_vt, const __grid_constant__ CUtensorMap tm_sfq,
                const __grid_constant__ CUtensorMap tm_sfk, const __grid_constant__ CUtensorMap tm_sfv, Fp4FwdParams p) {
  using L = Fp4Fwd2Smem<STAGES>;
  constexpr int D = kFp4D;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t q_full, kv_full[STAGES], kv_empty[STAGES], s_full, s_free, p_full[2], o_full[2];
  __shared__ uint64_t sc_full[2], sc_empty[2], o_ready[2];
  __shared__ float row_sc[2][128];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, q0 = blockIdx.x * 128;
  const int nk = p.Sk / 128, nst = 2 * nk;

  if (tid == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(&s_full, 1); mbar_init(&s_free, 4);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&p_full[b], 4); mbar_init(&o_full[b], 1);
      mbar_init(&sc_full[b], 4); mbar_init(&sc_empty[b], 4); mbar_init(&o_ready[b], 4);
    }
    fence_mbar_init();
  }
  if (warp == 9) tmem_alloc<256>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  constexpr uint32_t kP = 64, kSfQ = 80, kSfK = 88, kSfV = 104, kSfP = 120, kO = 128;

  if (warp < 4) {
    // =========================== softmax warps: thread = query row, 64 keys per step ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");           // 4 x 128 + 8 x 56 registers per lane = 960 = 12 x 80
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    const float c = p.sgq[bh] * p.sgk[bh] * p.qk_scale;
    const float2 c2 = make_float2(c, c);
    float m = -INFINITY, l = 0.f;
    for (int t = 0; t < nst; ++t) {
      const int pb = t & 1;
      mbar_wait(&s_full, pb);
      tc_fence_after();
      uint32_t r[64];
      tmem_ld64(lane_addr, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free);                        // the logits are in registers: S may be refilled
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 64; i += 8)
#pragma unroll
        for (int a = 0; a < 4; ++a) mx4[a] = fmaxf(mx4[a], fmaxf(__uint_as_float(r[i + 2 * a]), __uint_as_float(r[i + 2 * a + 1])));
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m, mx * c);
      const float resc = ex2_approx(m - m_new);
      m = m_new;
      if (t > 0) {                                                 // O *= 2^(m - m') is the correction warps' job
        const int k = t - 1, sl = k & 1;
        mbar_wait(&sc_empty[sl], ((k >> 1) & 1) ^ 1);
        row_sc[sl][row] = resc;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sc_full[sl]);
      }
      const float2 nm2 = make_float2(-m_new, -m_new);
      float2 ls2 = make_float2(0.f, 0.f);
      uint32_t pw[8], sfw = 0u;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float2 pe[8];
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(r[b * 16 + 2 * e]), __uint_as_float(r[b * 16 + 2 * e + 1])), c2, nm2);
          pe[e] = make_float2(ex2_approx(x.x), ex2_approx(x.y));
          if (e & 1) a1 = fmaxf(a1, fmaxf(pe[e].x, pe[e].y)); else a0 = fmaxf(a0, fmaxf(pe[e].x, pe[e].y));
          ls2 = __fadd2_rn(ls2, pe[e]);
        }
        const uint32_t sc = (uint32_t)__nv_cvt_float_to_fp8(fmaxf(a0, a1) * 448.0f, __NV_SATFINITE, __NV_E4M3);
        const float sf = fp4_e4m3_to_float(sc);
        const float inv = sf > 0.f ? __fdividef(2688.0f, sf) : 0.f;
        const float2 inv2 = make_float2(inv, inv);
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          float2 y[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = __fmul2_rn(pe[h8 * 4 + i], inv2);
          pw[2 * b + h8] = fp4_pack8(y);
        }
        sfw |= sc << (8 * b);
      }
      l = l * resc + (ls2.x + ls2.y);
      tmem_st8(lane_addr + kP + 8 * pb, pw);
      *reinterpret_cast<uint32_t*>(smem + L::off_sfp + pb * 512 + 16 * lane + 4 * warp) = sfw;
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[pb]);
    }
    // ---- epilogue: O * sgv / (2688 * l), log2-LSE
    mbar_wait(&o_full[(nst - 1) & 1], ((nst - 1) >> 1) & 1);
    tc_fence_after();
    const size_t gr = (size_t)bh * p.Sq + q0 + row;
    const float sc_o = __fdividef(p.sgv[bh], 2688.0f * l);
    __half* dst = p.O + gr * D;
#pragma unroll
    for (int ch = 0; ch < D / 32; ++ch) {
      uint32_t o[32];
      tmem_ld32(lane_addr + kO + ch * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 v;
        __half2 h2;
        h2 = __floats2half2_rn(__uint_as_float(o[i]) * sc_o, __uint_as_float(o[i + 1]) * sc_o); v.x = *reinterpret_cast<uint32_t*>(&h2);
        h2 = __floats2half2_rn(__uint_as_float(o[i + 2]) * sc_o, __uint_as_float(o[i + 3]) * sc_o); v.y = *reinterpret_cast<uint32_t*>(&h2);
        h2 = __floats2half2_rn(__uint_as_float(o[i + 4]) * sc_o, __uint_as_float(o[i + 5]) * sc_o); v.z = *reinterpret_cast<uint32_t*>(&h2);
        h2 = __floats2half2_rn(__uint_as_float(o[i + 6]) * sc_o, __uint_as_float(o[i + 7]) * sc_o); v.w = *reinterpret_cast<uint32_t*>(&h2);
        *reinterpret_cast<uint4*>(dst + ch * 32 + i) = v;
      }
    }
    p.lse[gr] = m + log2f(l);
  } else if (warp < 8) {
    // =========================== correction warps: O *= 2^(m - m') between P V(t-1) and P V(t) ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const int qd = warp & 3, row = qd * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(qd * 32) << 16);
    for (int t = 1; t < nst; ++t) {
      const int k = t - 1, sl = k & 1;
      mbar_wait(&sc_full[sl], (k >> 1) & 1);
      const float resc = row_sc[sl][row];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sc_empty[sl]);
      if (__any_sync(0xffffffffu, resc != 1.0f)) {
        mbar_wait(&o_full[(t - 1) & 1], ((t - 1) >> 1) & 1);
        tc_fence_after();
        const float2 rs2 = make_float2(resc, resc);
#pragma unroll
        for (int ch = 0; ch < D / 32; ++ch) {
          uint32_t o[32];
          tmem_ld32(lane_addr + kO + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 x = __fmul2_rn(make_float2(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), rs2);
            o[i] = __float_as_uint(x.x); o[i + 1] = __float_as_uint(x.y);
          }
          tmem_st32(lane_addr + kO + ch * 32, o);
        }
        tmem_st_wait();
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_ready[t & 1]);
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");            // warps 8-11 (10, 11 idle): one instruction for the warpgroup
    if (warp == 8) {
      // =========================== TMA producer ===========================
      if (elect_one()) {
        tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_vt);
        mbar_expect_tx(&q_full, L::kTile + L::kSf);
        tma_load_2d(smem + L::off_q, &tm_q, &q_full, 0, bh * p.Sq + q0);
        tma_load_2d(smem + L::off_sfq, &tm_sfq, &q_full, 0, (bh * p.Sq + q0) / 128);
        for (int j = 0; j < nk; ++j) {
          const int s = j % STAGES;
          mbar_wait(&kv_empty[s], ((j / STAGES) & 1) ^ 1);
          mbar_expect_tx(&kv_full[s], 2 * L::kTile + 2 * L::kSf);
          tma_load_2d(smem + L::off_k + s * L::kTile, &tm_k, &kv_full[s], 0, bh * p.Sk + j * 128);
          tma_load_2d(smem + L::off_sfk + s * L::kSf, &tm_sfk, &kv_full[s], 0, bh * (p.Sk / 128) + j);
          tma_load_2d(smem + L::off_v + s * L::kTile, &tm_vt, &kv_full[s], j * 64, bh * D);
          tma_load_2d(smem + L::off_sfv + s * L::kSf, &tm_sfv, &kv_full[s], 0, bh * (p.Sk / 128) + j);
        }
      }
    } else if (warp == 9) {
      // =========================== MMA issuer ===========================
      if (elect_one()) {
        constexpr uint32_t idesc_s = umma_idesc_bs(1, 1, 0, 0, 128, 64, 0);       // S step: N = 64 keys
        constexpr uint32_t idesc_o = umma_idesc_bs(1, 1, 0, 0, 128, 128, 0);      // O: N = D
        const uint32_t q_addr = smem_u32(smem + L::off_q);
        auto issue_s = [&](int t) {                                                // S = Q K_(step t)^T
          const int j = t >> 1, h = t & 1, s = j % STAGES, jb = j & 1;
          if (h == 0) {
            mbar_wait(&kv_full[s], (j / STAGES) & 1);
            tc_fence_after();
            const uint32_t sfk = smem_u32(smem + L::off_sfk + s * L::kSf);
            tmem_cp_sf(tbase + kSfK + jb * 8, sfk);
            tmem_cp_sf(tbase + kSfK + jb * 8 + 4, sfk + 512);
          }
          const uint32_t k_addr = smem_u32(smem + L::off_k + s * L::kTile) + h * (64 * 64);   // rows 64h .. 64h + 63
#pragma unroll
          for (int kd = 0; kd < 2; ++kd)
            umma_nvf4_ss(tbase, umma_smem_desc(q_addr + kd * 32, 16, 512, kSwz64), umma_smem_desc(k_addr + kd * 32, 16, 512, kSwz64), idesc_s,
                         tbase + kSfQ + kd * 4, tbase + kSfK + jb * 8 + kd * 4 + 2 * h, kd > 0);
          umma_commit(&s_full);
        };
        mbar_wait(&q_full, 0);
        tc_fence_after();
        tmem_cp_sf(tbase + kSfQ, smem_u32(smem + L::off_sfq));
        tmem_cp_sf(tbase + kSfQ + 4, smem_u32(smem + L::off_sfq) + 512);
        issue_s(0);
        for (int t = 0; t < nst; ++t) {
          const int j = t >> 1, h = t & 1, s = j % STAGES, jb = j & 1, pb = t & 1;
          if (t + 1 < nst) {
            mbar_wait(&s_free, pb);               // every softmax warp holds S(t) in registers
            tc_fence_after();
            issue_s(t + 1);
          }
          mbar_wait(&p_full[pb], (t >> 1) & 1);
          if (t > 0) mbar_wait(&o_ready[pb], ((t - 1) >> 1) & 1);
          tc_fence_after();
          if (h == 0) {
            const uint32_t sfv = smem_u32(smem + L::off_sfv + s * L::kSf);
            tmem_cp_sf(tbase + kSfV + jb * 8, sfv);
            tmem_cp_sf(tbase + kSfV + jb * 8 + 4, sfv + 512);
          }
          tmem_cp_sf(tbase + kSfP + pb * 4, smem_u32(smem + L::off_sfp + pb * 512));
          const uint32_t v_addr = smem_u32(smem + L::off_v + s * L::kTile) + h * 32;         // keys 64h .. of every D row
          umma_nvf4_ts(tbase + kO, tbase + kP + 8 * pb, umma_smem_desc(v_addr, 16, 512, kSwz64), idesc_o, tbase + kSfP + pb * 4,
                       tbase + kSfV + jb * 8 + 4 * h, t > 0);
          umma_commit(&o_full[pb]);
          if (h == 1) umma_commit(&kv_empty[s]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<256>(tbase);
}

}  // namespace qa

using namespace qa;

// q4, k4: [BH*S, 64] bytes (e2m1 pairs); vt4: [BH, 128, Sk/2] bytes; sf*: 512-byte atoms, 2 per 128-row tile; sg*: [BH] fp32
// (all produced by qa_fp4_quant_rows / qa_fp4_quant_vt).  O: fp16 [BH*Sq, 128]; lse: fp32 [BH*Sq] (log2 domain).
// variant 0 (default): one CTA per SM, de-phased exp warps, 128-key tiles; 1: two CTAs per SM, 64-key steps.
// flags: QA_FLAG_CAUSAL = the strict causal mask of the reference's baseline (key < query; row 0 of a head = average over all keys).
extern "C" int qa_fp4_fwd_ragged(const void* q4, const void* sfq, const void* sgq, const void* k4, const void* sfk, const void* sgk,
                                 const void* vt4, const void* sfv, const void* sgv, void* O_fp16, void* lse_f32, int BH, int Sq, int Sk,
                                 int Sk_valid, int D, int variant, int flags, float sm_scale, void* stream) {
  if (D != 128) return qa_fail(QA_ERR_SHAPE, "qa_fp4_fwd: D must be 128");
  if (BH <= 0 || Sq <= 0 || Sk <= 0 || Sq % 128 || Sk % 128) return qa_fail(QA_ERR_SHAPE, "qa_fp4_fwd: Sq, Sk must be positive multiples of 128");
  if ((long long)BH * Sq >= (1ll << 31) || (long long)BH * Sk >= (1ll << 31)) return qa_fail(QA_ERR_SHAPE, "qa_fp4_fwd: BH * S must stay below 2^31");
  const void* ptrs[11] = {q4, sfq, sgq, k4, sfk, sgk, vt4, sfv, sgv, O_fp16, lse_f32};
  for (int i = 0; i < 11; ++i) {
    if (!ptrs[i]) return qa_fail(QA_ERR_ALIGN, "qa_fp4_fwd: null pointer");
    if (i != 2 && i != 5 && i != 8 && i != 10 && ((uintptr_t)ptrs[i] & 15)) return qa_fail(QA_ERR_ALIGN, "qa_fp4_fwd: 16-byte alignment required");
  }
  if (variant < 0 || variant > 1) return qa_fail(QA_ERR_SHAPE, "qa_fp4_fwd: variant must be 0 (one CTA per SM, de-phased exp warps, 128-key tiles) or 1 (two CTAs per SM, 64-key steps)");
  if (Sk_valid <= Sk - 128 || Sk_valid > Sk) return qa_fail(QA_ERR_SHAPE, "qa_fp4_fwd: Sk_valid must lie in (Sk - 128, Sk]");
  if (Sk_valid != Sk && variant != 0) return qa_fail(QA_ERR_SHAPE, "qa_fp4_fwd: ragged sequences need variant 0");
  const bool causal = (flags & QA_FLAG_CAUSAL) != 0;
  if (flags & ~QA_FLAG_CAUSAL) return qa_fail(QA_ERR_SHAPE, "qa_fp4_fwd: unknown flag");
  if (causal && (Sq != Sk || variant != 0)) return qa_fail(QA_ERR_SHAPE, "qa_fp4_fwd: causal needs Sq == Sk and variant 0");
  constexpr int STAGES = 4;
  using L = Fp4FwdSmem<STAGES>;
  CUtensorMap tq, tk, tv, tsq, tsk, tsv;
  int rc;
  uint64_t dq[2] = {64, (uint64_t)BH * Sq}, dk[2] = {64, (uint64_t)BH * Sk}, dv[2] = {(uint64_t)Sk / 2, (uint64_t)BH * 128};
  uint64_t s64[1] = {64}, sv[1] = {(uint64_t)Sk / 2}, ssf[1] = {1024};
  uint32_t box[2] = {64, 128}, boxsf[2] = {256, 1};
  uint64_t dsq[2] = {256, (uint64_t)BH * Sq / 128}, dsk[2] = {256, (uint64_t)BH * Sk / 128};
  if ((rc = qa_make_tmap(&tq, q4, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dq, s64, box, 2))) return rc;
  if ((rc = qa_make_tmap(&tk, k4, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dk, s64, box, 2))) return rc;
  if ((rc = qa_make_tmap(&tv, vt4, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dv, sv, box, 2))) return rc;
  if ((rc = qa_make_tmap(&tsq, sfq, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dsq, ssf, boxsf, 0))) return rc;
  if ((rc = qa_make_tmap(&tsk, sfk, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dsk, ssf, boxsf, 0))) return rc;
  if ((rc = qa_make_tmap(&tsv, sfv, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dsk, ssf, boxsf, 0))) return rc;
  Fp4FwdParams p;
  p.sgq = (const float*)sgq; p.sgk = (const float*)sgk; p.sgv = (const float*)sgv;
  p.O = (__half*)O_fp16; p.lse = (float*)lse_f32; p.Sq = Sq; p.Sk = Sk; p.BH = BH; p.Sk_valid = Sk_valid;
  // sm_scale <= 0: 1 / sqrt(D).  (A head dimension of 64 runs as D = 128 with zero-padded columns and sm_scale = 1/8.)
  p.qk_scale = sm_scale > 0.f ? sm_scale * 1.44269504f : (float)((1.0 / sqrt((double)D)) * 1.44269504);
  if (variant == 1) {
    using L2 = Fp4Fwd2Smem<3>;
    auto kern2 = fp4_fwd2_kernel<3>;
    cudaError_t e2 = cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, L2::total);
    if (e2 != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e2));
    kern2<<<dim3(Sq / 128, BH), 384, L2::total, (cudaStream_t)stream>>>(tq, tk, tv, tsq, tsk, tsv, p);
    return qa_check_launch("qa_fp4_fwd");
  }
  auto kern = causal ? fp4_fwd_kernel<STAGES, true> : fp4_fwd_kernel<STAGES, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  kern<<<causal ? dim3((unsigned)((size_t)BH * (Sq / 128))) : dim3(Sq / 128, BH), 640, L::total, (cudaStream_t)stream>>>(tq, tk, tv, tsq, tsk, tsv, p);
  if (causal) {
    int rc2 = qa_check_launch("qa_fp4_fwd");
    if (rc2) return rc2;
    fp4_row0_fixup_kernel<<<BH, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)vt4, (const uint8_t*)sfv, p.sgv, p.O, p.lse, Sk, Sk_valid);
  }
  return qa_check_launch("qa_fp4_fwd");
}

extern "C" int qa_fp4_fwd(const void* q4, const void* sfq, const void* sgq, const void* k4, const void* sfk, const void* sgk,
                          const void* vt4, const void* sfv, const void* sgv, void* O_fp16, void* lse_f32, int BH, int Sq, int Sk,
                          int D, int variant, int flags, void* stream) {
  return qa_fp4_fwd_ragged(q4, sfq, sgq, k4, sfk, sgk, vt4, sfv, sgv, O_fp16, lse_f32, BH, Sq, Sk, Sk, D, variant, flags, 0.f, stream);
}
