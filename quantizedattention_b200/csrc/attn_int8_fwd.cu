// int8 SageAttention3-style forward (reference attention_int8.py:170-257) for sm_100a.
//
// One CTA = one 128-row query tile of one (batch, head).
// NSPLIT == 2 (the tuned 128-key tile; 20 warps): the online softmax is a two-stage pipeline across warpgroups
//   exp warps        0..7    second stage: fp16 logits (TMEM) -> exp2, row sums, int8 P -> TMEM; two warps per 32 rows,
//                            64 columns each
//   logit warps      8..11   first stage, whole rows: TMEM int32 S -> packed fp16 logits written back over S, running
//                            maximum, rescale factor and P scale of the tile, published per 32-row group
//   drain warps      12..15  TMEM int32 P.V partial -> fp32 O accumulators (all D columns) in registers
//   TMA producer 16, MMA issuer 17 (tcgen05.mma kind::i8; S = Q K^T from shared memory, Opart = P V with P from TMEM),
//   18..19 idle; setmaxnreg moves the registers of everybody else to the drain warps (80/80/80/192/40 per sub-partition).
//   TMEM (512 cols): S[3] at 0/128/256 (Q K^T runs two tiles ahead), one Opart at 384.
// NSPLIT == 1 (Bkv = 32 / 64 and the one-thread-per-row variant): softmax warps [0,4) do both stages, correction warps
//   [4,8), producer 8, MMA 9; TMEM S[2] at 0/128, Opart[2] at 256/384.
// The int32 P.V accumulator cannot span k-tiles (the P scale is per row per k-tile, the V scale per k-tile), so each
// k-tile's partial is drained to registers.
// Numerics follow the reference step by step (fp16 logits, fp16 running max, fp16 subtraction, per-row P scale
// exp2(rowmax - m)/127, truncation toward zero); see DESIGN.md for the two tolerance-level deviations
// (single fused scale multiply; reciprocal multiply instead of divide for P/sp).
#include "qa_ptx.cuh"
#include "qa_host.h"
#include <type_traits>

namespace qa {

constexpr int kBM = 128;    // query rows per CTA (= tcgen05 M)
#ifndef QA_DRAIN_W
#define QA_DRAIN_W 16
#endif
#ifndef QA_ROLE_PERM
#define QA_ROLE_PERM 1
#endif
// BN = keys per k-tile = Bkv (the reference's tunable, attention_int8.py:158): 128 is the tuned value; 32 (the reference
// default) and 64 run the same kernel with narrower S tiles -- one online-softmax step, one P scale per row and one
// drained P.V partial per Bkv keys, exactly as the reference's k-tile loop -- at proportionally more TMEM drains.

template <int D, int NSPLIT, int STAGES, int BN>
struct Int8FwdSmem {
  static constexpr int kQBytes = kBM * D;
  static constexpr int kKBytes = BN * D;
  static constexpr int kVBytes = BN * D;
  static constexpr int off_q = 0;
  static constexpr int off_k = off_q + kQBytes;
  static constexpr int off_v = off_k + STAGES * kKBytes;
  static constexpr int off_c = off_v + STAGES * kVBytes;      // constant fp16 tiles (A, B) of the accumulator-initialising MMA (MG):
  static constexpr int off_end = off_c + 2 * 1024;            // one 8-row swizzle atom each, shared by all row groups (SBO = 0)
  static constexpr int total = off_end + 1024;   // + alignment slack
};

struct Int8FwdParams {
  const __half* sq;      // [BH*Sq/Bq]
  const __half* sk;      // [BH*Sk/128]
  const __half* sv;      // [BH*Sk/128]
  __half* O;             // [BH*Sq, D] fp16
  __half* lse16;         // [BH*Sq]
  float* lse32;          // [BH*Sq] (may be null)
  float* m_out;          // optional state outputs for ring attention: unnormalised O + (m, l); null otherwise
  float* l_out;
  float* O_acc_out;
  const float* m_in;     // optional running state to continue from (previous K/V shards of the ring); null = fresh
  const float* l_in;
  const float* O_acc_in;
  int Sq, Sk, Bq;
  int Sk_valid;          // keys [Sk_valid, Sk) of every head are padding (ragged sequence, hl.tile clamps the last tile: attention_int8.py:170,176): weight exactly 0
  float qk_scale;
  long long* dbg;        // optional timeline buffer [tile][16] of SM clock stamps written by CTA (0,0) (tools/timeline.py)
};

// Timeline stamps exist only in development builds (-DQA_DEV_TIMELINE, libqattn_dev.so; tools/timeline.py): the shipped
// kernel carries no debug hooks.
#ifdef QA_DEV_TIMELINE
#define QA_TL(slot)                                                                                   \
  do {                                                                                                \
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && j < 64)                \
      p.dbg[j * 16 + (slot)] = clock64();                                                             \
  } while (0)
#define QA_TLX(cond, idx, slot)                                                                       \
  do {                                                                                                \
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && (cond)) p.dbg[(idx) * 16 + (slot)] = clock64(); \
  } while (0)
#else
#define QA_TL(slot) do { } while (0)
#define QA_TLX(cond, idx, slot) do { } while (0)
#endif

// MG ("magic" accumulators, two-stage kernel only): every S and P.V accumulator is initialised by a kind::f16 MMA over a
// constant tile to kMagic (qa_ptx.cuh), so the int32 results are read back as floats kMagic + x and the int -> float
// conversions fold into the scale FMAs; the exp warps fold the P scale into the exponent (one FHADD per element takes the
// fp16 logit to fp32, subtracts the running maximum and adds log2(127 / sp)), so exp2 yields P / sp directly.
// NDR (two-stage kernel): drain warps per 32-row group.  NDR == 2: 24 warps, each drain warp owns half of the D output
// columns, so the drain of tile j (the only consumer of the single P.V partial buffer) takes half as long and the next
// P V can start that much earlier; registers per role 72 (exp, logit) / 112 (drain) / 40 <= 6 x 80 per sub-partition.
template <int NSPLIT, int NDR>
struct Int8FwdRoles {
  static constexpr int kThreads = NSPLIT == 1 ? 384 : 512 + 128 * NDR;
  static constexpr int kDrain0 = NSPLIT == 1 ? 4 : 12;           // first drain (correction) warp
  static constexpr int kNumDrain = NSPLIT == 1 ? 4 : 4 * NDR;
  static constexpr int kTmaWarp = kDrain0 + kNumDrain, kMmaWarp = kTmaWarp + 1;
};

// FP8 (only with the MG kernel structure): Q, K, V and P are e4m3 (scale = amax / 448 instead of / 127,
// round to nearest), the contractions are tcgen05 kind::f8f6f4 with fp32 accumulation, so the accumulators are floats
// already: no initialising MMA, no magic constant; everything else (fp16 logits, fp16 running maximum, per-row P scale per
// k-tile, per-tile V scale, fp32 O in registers) is the int8 pipeline.
template <int D, int NSPLIT, int STAGES, int BN, bool RN, bool CAUSAL, bool MG, int NDR, bool FP8>
__global__ void __launch_bounds__((Int8FwdRoles<NSPLIT, NDR>::kThreads), 1)
int8_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, Int8FwdParams p) {
  using L = Int8FwdSmem<D, NSPLIT, STAGES, BN>;
  constexpr int kBN = BN;
  constexpr int NC = kBN / NSPLIT;       // S columns per softmax thread
  static_assert(NC % 32 == 0, "a softmax thread handles a multiple of 32 columns");
  constexpr int DC = D / NSPLIT;         // O columns per correction thread
  constexpr int kSoftWarps = 4 * NSPLIT;
  using R = Int8FwdRoles<NSPLIT, NDR>;
  static_assert(NDR == 1 || NSPLIT == 2, "two drain warps per row group exist in the two-stage kernel only");
  // NSPLIT == 2: S[b] columns hold, per 64-column half h, the packed fp16 logits at [64h, 64h+32); P (int8) at [32, 64)
  constexpr uint32_t kPOff = (NSPLIT == 2) ? 32 : 0;
  // TMEM buffers: NSPLIT == 1: S[2] at 0/128, Opart[2] at 256/384.  NSPLIT == 2: S[3] at 0/128/256 (Q K^T runs two tiles
  // ahead, so the logit warps never wait for the tensor pipe behind a P V) and a single Opart at 384 (the drain of tile
  // j is shorter than a tile period, P V of tile j+1 waits for it).
  // BN == 256 (Bkv = 256, the reference tunable's upper end): one 256-column S buffer, the logits are recomputed in pass 2
  constexpr int kSBuf = (NSPLIT == 2) ? 3 : (BN > 128 ? 1 : 2), kOBuf = (NSPLIT == 2) ? 1 : 2;
  constexpr int kSStride = BN > 128 ? BN : 128;                // TMEM columns between S buffers
  constexpr uint32_t kOCol = kSBuf * kSStride;
  constexpr bool kKeep = (NC <= 128);                          // keep the packed fp16 logits in registers between the passes
  static_assert(NSPLIT == 1 || BN == 128, "the two-stage softmax is laid out for 128-key tiles");
  static_assert(!MG || NSPLIT == 2, "magic accumulators are built into the two-stage kernel");
  static_assert(!FP8 || (MG && NDR == 1 && !CAUSAL), "the fp8 forward reuses the structure of the magic two-stage kernel");
  constexpr bool kInit = MG && !FP8;                           // accumulators initialised to kMagic by a kind::f16 MMA
  constexpr float kQMax = FP8 ? 448.0f : 127.0f;               // largest quantised magnitude of P
  constexpr uint32_t kLayoutQK = (D == 128) ? kSwz128 : kSwz64;   // rows of D bytes
  constexpr uint32_t kSboQK = (D == 128) ? 1024 : 512;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t q_full, k_full[STAGES], k_empty[STAGES], v_full[STAGES], v_empty[STAGES];
  __shared__ uint64_t s_full[3], p_full[3], o_full[2], o_empty[2], sc_full[2], sc_empty[2], fin_full;
  // NSPLIT == 2: logits + per-row parameters (rescale, 127/sp, m, sp*sv) of tile j published in ring slot j & 3: slot
  // j is rewritten for tile j+4, i.e. after P V of tile j+2, which itself waits for the drain of tile j
  __shared__ uint64_t lg_full[4][4];
  __shared__ float4 prm_s[4][kBM];
  __shared__ __half prm_m[4][kBM];           // running maximum of the tile (MG: the exp warps subtract it in fp16, as the reference)
  __shared__ uint32_t tmem_base_s;
  __shared__ float2 row_sc[2][kBM];          // per tile parity: (rescale, sp*sv) per row
  __shared__ float l_part[2][kBM];
  __shared__ __half m_fin[kBM];

  // Role of a warp.  The hardware scheduler prefers the highest warp id among the eligible warps of a sub-partition, so
  // the two-stage kernel gives the high ids to the stage that binds (the exp warps: 16 exp2 per clock and SM), then the
  // logit warps, then the drain warps; `warp` below is the LOGICAL id (exp 0-7, logit 8-11, drain 12-15, TMA 16, MMA 17).
  const int tid = threadIdx.x, lane = tid & 31, pwarp = tid >> 5;
  const int warp = (NSPLIT == 2 && NDR == 1 && QA_ROLE_PERM)
                       ? (pwarp >= 8 && pwarp < 16 ? pwarp - 8 : pwarp >= 4 && pwarp < 8 ? pwarp + 4 : pwarp < 4 ? pwarp + 12
                          : pwarp == 18 ? 16 : pwarp == 19 ? 17 : pwarp + 2)
                       : pwarp;
  // CAUSAL (strict mask, key < query): a query tile visits the k-tiles up to its own; heaviest tiles first
  const int bh = blockIdx.y, q0 = (CAUSAL ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * kBM;
  const int nk = CAUSAL ? min(p.Sk / kBN, q0 / kBN + 1) : (p.Sk_valid + kBN - 1) / kBN;   // tiles beyond the last valid key are skipped
  const int ktail = p.Sk_valid - (nk - 1) * kBN;               // valid keys of the last tile (== kBN unless the sequence is ragged)

  if (tid == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], MG ? 4 : kSoftWarps);             // MG: one exp warp per 32-row group works on a given tile
      if (b == 0) { mbar_init(&s_full[2], 1); mbar_init(&p_full[2], MG ? 4 : kSoftWarps); }
      for (int qd = 0; qd < 4; ++qd) { mbar_init(&lg_full[b][qd], 1); mbar_init(&lg_full[b + 2][qd], 1); }
      mbar_init(&o_full[b], 1); mbar_init(&o_empty[b], NSPLIT == 2 ? R::kNumDrain : 4);   // every drain warp of the tile
      mbar_init(&sc_full[b], 4); mbar_init(&sc_empty[b], kSoftWarps);
    }
    mbar_init(&fin_full, kSoftWarps + (NSPLIT == 2 ? 4 : 0));
    fence_mbar_init();
  }
  if (warp == R::kMmaWarp) tmem_alloc<512>(&tmem_base_s);
  if (kInit) {                                                   // constant operand tiles: every fp16 element = 1024 (A) / 768 (B)
    const uint32_t c_addr = smem_u32(smem + L::off_c);
    if (tid < 64) sts128(c_addr + tid * 16, kMagicElemA2, kMagicElemA2, kMagicElemA2, kMagicElemA2);
    else if (tid < 128) sts128(c_addr + tid * 16, kMagicElemB2, kMagicElemB2, kMagicElemB2, kMagicElemB2);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  // NSPLIT == 2 (20 warps, 5 per SM sub-partition, 96 registers each at launch = the CTA's whole budget): the exp warps,
  // the logit warps and the last warpgroup (TMA, MMA, two idle warps) hand registers to the drain warps, which hold the
  // fp32 O accumulators of all D columns: 80 + 80 + 80 + 192 + 40 <= 5 x 96 per lane of a sub-partition (setmaxnreg.inc
  // can only take what the CTA's other warps released)
  if (warp < kSoftWarps) {
    // =========================== softmax warps ===========================
    if (NSPLIT == 2) { if (NDR == 2 || MG) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;"); else asm volatile("setmaxnreg.dec.sync.aligned.u32 80;"); }
    const int split = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const int c0 = split * NC;
    const float sq_f = __half2float(p.sq[((size_t)bh * p.Sq + q0 + row) / p.Bq]);
    __half m16 = __float2half_rn(-INFINITY);
    float l = (split == 0) ? 1.0f : 0.0f;          // reference initialises l to 1.0 (attention_int8.py:173)
    if (p.m_in != nullptr) {                       // ring: continue the online softmax of earlier K/V shards
      const size_t gr = (size_t)bh * p.Sq + q0 + row;
      m16 = __float2half_rn(p.m_in[gr]);
      l = (split == 0) ? p.l_in[gr] : 0.0f;
    }
    if (MG) {
      // ---- exp warps of the magic kernel.  The two exp warps of a 32-row group are DE-PHASED: warp e = split handles the
      //      tiles j = e (mod 2), all 128 columns, so that one of them is in its MUFU-bound inner loop while the other
      //      waits for logits / loads them / hands P over: the XU (16 exp2 per clock and SM) is the binding unit of this
      //      stage and stays busy.  y = P / sp = exp2(fp16(S16 - m) + log2(127 / sp_e)): the subtraction is the reference's
      //      fp16 one (attention_int8.py:211-213), prm.y = log2(127) - (rmax - m) comes from the logit warps, the conversion to
      //      fp32 and the addition are one FHADD.  trunc(y) is the bit pattern of the subnormal y * 2^-149 rounded toward
      //      zero; the I2IP pack saturates y in (127, 128.5) to 127.
      const int e = split, qd = warp & 3;
      const float2 tiny2 = make_float2(1.401298464324817e-45f, 1.401298464324817e-45f);
      for (int j = e; j < nk; j += 2) {
        const int sb = j % kSBuf;
        if (warp == 0) QA_TL(0);
        mbar_wait(&lg_full[j & 3][qd], (j >> 2) & 1);
        tc_fence_after();
        if (warp == 0) QA_TL(1);
        const float4 prm = prm_s[j & 3][row];
        float resc = prm.x;
        if (j >= 2) resc *= prm_s[(j - 1) & 3][row].x;             // the tile the other exp warp of this row group handled
        const float kk = prm.y;
        const __half2 m2 = __half2half2(prm_m[j & 3][row]);
        float2 ys2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                           // 64 columns at a time
          uint32_t lg[32];
          tmem_ld32(lane_addr + sb * 128 + hf * 64, lg);
          tmem_ld_wait();
          if (warp == 0 && hf == 0) QA_TL(2);
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint32_t w[8];
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              const __half2 d0 = __hsub2(*reinterpret_cast<const __half2*>(&lg[g * 16 + q4 * 2]), m2);       // fp16 subtraction (:211-213)
              const __half2 d1 = __hsub2(*reinterpret_cast<const __half2*>(&lg[g * 16 + q4 * 2 + 1]), m2);
              const uint32_t h0 = *reinterpret_cast<const uint32_t*>(&d0), h1 = *reinterpret_cast<const uint32_t*>(&d1);
              const float2 y0 = make_float2(ex2_approx(fhadd_lo(h0, kk)), ex2_approx(fhadd_hi(h0, kk)));
              const float2 y1 = make_float2(ex2_approx(fhadd_lo(h1, kk)), ex2_approx(fhadd_hi(h1, kk)));
              ys2 = __fadd2_rn(ys2, y0);
              ys2 = __fadd2_rn(ys2, y1);
              if (FP8) {                                             // y in [0, 448] -> e4m3, round to nearest, saturating
                w[q4] = cvt_e4m3x2(y0.x, y0.y) | (cvt_e4m3x2(y1.x, y1.y) << 16);
              } else {
                const float2 q0 = RN ? __fmul2_rn(y0, tiny2) : __fmul2_rz(y0, tiny2);
                const float2 q1 = RN ? __fmul2_rn(y1, tiny2) : __fmul2_rz(y1, tiny2);
                w[q4] = pack_sat_s8x4(__float_as_int(q0.x), __float_as_int(q0.y), __float_as_int(q1.x), __float_as_int(q1.y));
              }
            }
            tmem_st8(lane_addr + sb * 128 + kPOff + hf * 16 + g * 8, w);
          }
        }
        tmem_st_wait();
        l = l * resc + (ys2.x + ys2.y) * prm.z;                    // sum(P) = sum(y) * sp   (attention_int8.py:215-223)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[sb]);
        if (warp == 0) QA_TL(5);
      }
      if (nk > 0 && ((nk - 1) & 1) != e) {                         // the last tile belonged to the other warp: its rescale
        mbar_wait(&lg_full[(nk - 1) & 3][qd], ((nk - 1) >> 2) & 1);
        l *= prm_s[(nk - 1) & 3][row].x;
      }
    } else
    for (int j = 0; j < nk; ++j) {
      const int b = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int sb = j % kSBuf;                                    // S buffer of this tile
      const float sk_f = __half2float(p.sk[((size_t)bh * p.Sk) / kBN + j]);
      const float sv_f = __half2float(p.sv[((size_t)bh * p.Sk) / kBN + j]);
      const float c = sq_f * sk_f * p.qk_scale;
      const float2 c2 = make_float2(c, c);
      if (warp == 0) QA_TL(0);
      __half2 sh[kKeep ? NC / 2 : 1];
      __half rmax = __float2half_rn(0.f);
      float4 prm = make_float4(0.f, 0.f, 0.f, 0.f);
      if (NSPLIT == 2) {
        // ---- two-stage softmax: the logit warps (below) have turned S into packed fp16 logits in TMEM and published
        //      the row maxima; this warp only does the exponential / quantisation half of the work
        mbar_wait(&lg_full[j & 3][warp & 3], (j >> 2) & 1);
        tc_fence_after();
        if (warp == 0) QA_TL(1);
        uint32_t lg[32];
        tmem_ld32(lane_addr + sb * 128 + split * 64, lg);
        prm = prm_s[j & 3][row];
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) sh[i] = *reinterpret_cast<__half2*>(&lg[i]);
        if (warp == 0) QA_TL(2);
      } else {
        mbar_wait(&s_full[sb], (j / kSBuf) & 1);
        tc_fence_after();
        if (warp == 0) QA_TL(1);
        // ---- pass 1: int32 -> fp16 logits (packed), row max
        __half2 mx2 = __float2half2_rn(-INFINITY);
#pragma unroll
        for (int ch = 0; ch < NC / 32; ++ch) {
          uint32_t r[32];
          tmem_ld32(lane_addr + sb * kSStride + c0 + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {                          // packed fp32x2 multiply (FMUL2): half the issue slots
            const float2 a = __fmul2_rn(make_float2(__int2float_rn((int)r[2 * i]), __int2float_rn((int)r[2 * i + 1])), c2);
            __half2 h = __float22half2_rn(a);
            if (j == nk - 1 && ktail < kBN) {                      // padding keys of a ragged sequence
              const int col = c0 + ch * 32 + 2 * i;
              const __half ninf = __float2half_rn(-INFINITY);
              if (col >= ktail) h = __halves2half2(ninf, __high2half(h));
              if (col + 1 >= ktail) h = __halves2half2(__low2half(h), ninf);
            }
            if (kKeep) sh[ch * 16 + i] = h;
            mx2 = __hmax2(mx2, h);
          }
        }
        rmax = __hmax(__low2half(mx2), __high2half(mx2));
        if (warp == 0) QA_TL(2);
      }
      __half m_new;
      float rescale, inv_sp;
      if (NSPLIT == 2) {                             // the logit warps own the running maximum and the per-row scales
        rescale = prm.x; inv_sp = prm.y; m_new = __float2half_rn(prm.z);
      } else {
        m_new = __hmax(m16, rmax);
        rescale = ex2_approx(__half2float(__hsub(m16, m_new)));                 // fp16 subtraction (:217-219)
        const float sp_e = ex2_approx(__half2float(__hsub(rmax, m_new)));       // (:232-234) sp = sp_e / 127
        inv_sp = __fdividef(127.0f, sp_e);
        m16 = m_new;
        // ---- hand (rescale, sp*sv) to the correction warps
        mbar_wait(&sc_empty[b], ph ^ 1);
        row_sc[b][row] = make_float2(rescale, sp_e * (1.0f / 127.0f) * sv_f);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sc_full[b]);
      }
      // ---- pass 2: P = exp2(S16 - m), l += sum(P), P_i8 = trunc(P / sp) -> written back to TMEM over the S columns
      //      (4 int8 per column): the P V MMA takes its A operand straight from TMEM (no shared-memory round trip,
      //      no proxy fence, no second buffer to wait for - the S buffer is ours until that MMA has been issued)
      if (warp == 0) { QA_TL(3); QA_TL(4); }
      const __half2 m2 = __half2half2(m_new);       // (causal: the logit warps publish 0 while a row has seen no key)
      float2 ls2 = make_float2(0.f, 0.f);
      const float2 inv2 = make_float2(inv_sp, inv_sp), magic2 = make_float2(8388608.0f, 8388608.0f);
      {                                                            // RN: the rounding mode is an FFMA2 modifier
#pragma unroll
        for (int g = 0; g < NC / 32; ++g) {
          uint32_t w[8];
          uint32_t rr[32];
          if (!kKeep) {                                            // wide tile: S is read again, the logits recomputed
            tmem_ld32(lane_addr + sb * kSStride + c0 + g * 32, rr);
            tmem_ld_wait();
          }
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            uint32_t bytes[4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int e2 = q4 * 2 + h2;                          // pair of columns inside the 32-column group
              __half2 hl = kKeep ? sh[kKeep ? g * 16 + e2 : 0]
                                 : __float22half2_rn(__fmul2_rn(make_float2(__int2float_rn((int)rr[2 * e2]), __int2float_rn((int)rr[2 * e2 + 1])), c2));
              if (!kKeep && j == nk - 1 && ktail < kBN) {
                const int col = c0 + g * 32 + 2 * e2;
                const __half ninf = __float2half_rn(-INFINITY);
                if (col >= ktail) hl = __halves2half2(ninf, __high2half(hl));
                if (col + 1 >= ktail) hl = __halves2half2(__low2half(hl), ninf);
              }
              const float2 f = __half22float2(__hsub2(hl, m2));   // fp16 subtraction (:211-213)
              const float2 pp = make_float2(ex2_approx(f.x), ex2_approx(f.y));
              ls2 = __fadd2_rn(ls2, pp);
              // low byte of the biased sum = trunc(P/sp) (reference) or its nearest-even rounding (accuracy mode)
              const float2 qf = RN ? __ffma2_rn(pp, inv2, magic2) : __ffma2_rz(pp, inv2, magic2);
              bytes[h2 * 2] = __float_as_uint(qf.x);
              bytes[h2 * 2 + 1] = __float_as_uint(qf.y);
            }
            w[q4] = pack_low_bytes(bytes[0], bytes[1], bytes[2], bytes[3]);
          }
          tmem_st8(lane_addr + sb * kSStride + kPOff + c0 / 4 + g * 8, w);
        }
      }
      tmem_st_wait();
      l = l * rescale + (ls2.x + ls2.y);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
      if (warp == 0) QA_TL(5);
    }
    l_part[split][row] = l;
    if (NSPLIT == 1) m_fin[row] = m16;
    __syncwarp();
    if (lane == 0) mbar_arrive(&fin_full);
  } else if (NSPLIT == 2 && warp < kSoftWarps + 4) {
    // =========================== logit warps (two-stage softmax, NSPLIT == 2) ===========================
    // First stage of the online softmax for one 32-row group, whole rows (no cross-warp exchange): int32 S -> packed
    // fp16 logits written back over the first half of each 64-column half of S[b], running maximum, rescale factor
    // and P scale of the tile published for the exp warps (second stage) and the drain warp.
    if (NDR == 2 || MG) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;"); else asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(qd * 32) << 16);
    const float sq_f = __half2float(p.sq[((size_t)bh * p.Sq + q0 + row) / p.Bq]);
    const __half* sk_p = p.sk + ((size_t)bh * p.Sk) / kBN;
    const __half* sv_p = p.sv + ((size_t)bh * p.Sk) / kBN;
    __half m16 = __float2half_rn(-INFINITY);
    if (p.m_in != nullptr) m16 = __float2half_rn(p.m_in[(size_t)bh * p.Sq + q0 + row]);
    for (int j = 0; j < nk; ++j) {
      const int sb = j % kSBuf;
      const float c = MG ? magic_scale(sq_f * __half2float(sk_p[j]) * p.qk_scale) : sq_f * __half2float(sk_p[j]) * p.qk_scale;
      const float2 c2 = make_float2(c, c), nb2 = make_float2(-kMagic * c, -kMagic * c);       // kMagic * c is exact
      if (qd == 0) QA_TL(12);
      mbar_wait(&s_full[sb], (j / kSBuf) & 1);
      tc_fence_after();
      if (qd == 0) QA_TL(13);
      __half2 mx2 = __float2half2_rn(-INFINITY);
      auto pass1 = [&](auto masked, auto tail) {                   // masked: the diagonal tile of a causal head; tail: ragged last tile
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {                          // 32 columns in, 16 columns (32 packed pairs) out
          const uint32_t src = lane_addr + sb * 128 + ch * 32;
          const uint32_t dst = lane_addr + sb * 128 + (ch >> 1) * 64 + (ch & 1) * 16;   // read before it is overwritten
          uint32_t r[32];
          tmem_ld32(src, r);
          tmem_ld_wait();
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {                          // packed fp32x2 multiply (FMUL2): half the issue slots
            const float2 a = FP8 ? __fmul2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2)        // fp32 accumulator
                             : MG ? __ffma2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, nb2)   // (kMagic + x) * c - kMagic * c
                                  : __fmul2_rn(make_float2(__int2float_rn((int)r[2 * i]), __int2float_rn((int)r[2 * i + 1])), c2);
            __half2 h = __float22half2_rn(a);
            if (decltype(masked)::value) {                         // strict causal: keep key < query (same tile: col < row)
              const int col = ch * 32 + 2 * i;
              const __half ninf = __float2half_rn(-INFINITY);
              if (col >= row) h = __halves2half2(ninf, __high2half(h));
              if (col + 1 >= row) h = __halves2half2(__low2half(h), ninf);
            }
            if (decltype(tail)::value) {                           // padding keys of a ragged sequence
              const int col = ch * 32 + 2 * i;
              const __half ninf = __float2half_rn(-INFINITY);
              if (col >= ktail) h = __halves2half2(ninf, __high2half(h));
              if (col + 1 >= ktail) h = __halves2half2(__low2half(h), ninf);
            }
            mx2 = __hmax2(mx2, h);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
          }
          tmem_st16(dst, w);
        }
      };
      if (CAUSAL && j * kBN == q0) pass1(std::true_type{}, std::false_type{});
      else if (!CAUSAL && j == nk - 1 && ktail < kBN) pass1(std::false_type{}, std::true_type{});
      else pass1(std::false_type{}, std::false_type{});
      const __half rmax = __hmax(__low2half(mx2), __high2half(mx2));
      const __half m_new = __hmax(m16, rmax);
      float rescale = ex2_approx(__half2float(__hsub(m16, m_new)));             // fp16 subtraction (:217-219)
      float sp_e = ex2_approx(__half2float(__hsub(rmax, m_new)));               // (:232-234) sp = sp_e / 127
      float inv_sp = __fdividef(127.0f, sp_e);
      if (CAUSAL) {                                   // a row may have no visible key in this tile / so far: (-inf) - (-inf)
        if (__hisinf(m_new)) rescale = 1.0f;
        if (__hisinf(rmax)) { sp_e = 0.f; inv_sp = 0.f; }
      }
      m16 = m_new;
      if (MG) {                                       // (rescale, exponent offset of y = P / sp, sp, sp * sv)
        const float dsp = __half2float(__hsub(rmax, m_new));
        // log2(127) + eps (fp8: log2(448)): the row maximum quantises to the largest code
        const float kk = (CAUSAL && __hisinf(rmax)) ? 0.f : (FP8 ? 8.807354922057604f : 6.988684686772166f + 2.0e-6f) - dsp;
        prm_m[j & 3][row] = (CAUSAL && __hisinf(m_new)) ? __float2half_rn(0.f) : m_new;
        prm_s[j & 3][row] = make_float4(rescale, kk, sp_e * (1.0f / kQMax), sp_e * (1.0f / kQMax) * __half2float(sv_p[j]));
      } else
      prm_s[j & 3][row] = make_float4(rescale, inv_sp, (CAUSAL && __hisinf(m_new)) ? 0.f : __half2float(m_new),
                                      sp_e * (1.0f / 127.0f) * __half2float(sv_p[j]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&lg_full[j & 3][qd]);
      if (qd == 0) QA_TL(14);
    }
    m_fin[row] = m16;
    __syncwarp();
    if (lane == 0) mbar_arrive(&fin_full);
  } else if (warp < R::kTmaWarp) {
    // =========================== drain (correction) warps ===========================
    // NSPLIT == 1: 4 warps, scales handed over by the softmax warps.  NSPLIT == 2: warps 12..15, one per 32-row group,
    // all D output columns (the registers come from the other warpgroups), scales published by the logit warps.
    constexpr bool kTwoStage = (NSPLIT == 2);
    constexpr int DCx = kTwoStage ? D / NDR : DC;
    // registers per sub-partition: NDR == 1: 80 + 80 + 80 + 192 + 40 (MG: 72 + 72 + 72 + 224 + 40) <= 5 x 96; NDR == 2: 4 x 72 + 2 x 112 ... <= 6 x 80
    if (kTwoStage) {
      if (NDR == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
      else if (MG) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
      else asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
    }
    const int cw = warp - R::kDrain0;
    const int split = cw >> 2;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(qd * 32) << 16);
    const int d0 = split * DCx;
    float2 acc2[DCx / 2];                                  // fp32x2 accumulators: one FFMA2 per two elements
#pragma unroll
    for (int i = 0; i < DCx / 2; ++i) acc2[i] = make_float2(0.f, 0.f);
    if (p.O_acc_in != nullptr) {
      const float* src = p.O_acc_in + ((size_t)bh * p.Sq + q0 + row) * D + d0;
#pragma unroll
      for (int i = 0; i < DCx; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(src + i);
        acc2[i / 2] = make_float2(t.x, t.y); acc2[i / 2 + 1] = make_float2(t.z, t.w);
      }
    }
    float s_pend = 1.0f, bias = 0.f;
    for (int j = 0; j < nk; ++j) {
      const int b = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int ob = j % kOBuf;                                    // P.V partial buffer of this tile
      float2 sc;
      if (kTwoStage) {
        mbar_wait(&lg_full[j & 3][qd], (j >> 2) & 1);
        const float4 prm = prm_s[j & 3][row];
        sc = make_float2(prm.x, prm.w);
      } else {
        mbar_wait(&sc_full[b], ph);
        sc = row_sc[b][row];
        __syncwarp();
        if (lane == 0) mbar_arrive(&sc_empty[b]);
      }
      // Lazy rescale: the accumulator holds O / s_pend, so O*rescale + x*c (attention_int8.py:225, 249-250) costs one
      // FMA per element: s_pend *= rescale; acc += x * (c / s_pend).
      float fold = 1.0f;                            // usually nothing to do
      if (sc.x != 0.f) s_pend *= sc.x; else { fold = 0.f; s_pend = 1.0f; }
      if (s_pend < 1e-18f) { fold = s_pend; s_pend = 1.0f; }
      if (__any_sync(0xffffffffu, fold != 1.0f)) {
#pragma unroll
        for (int i = 0; i < DCx / 2; ++i) acc2[i] = __fmul2_rn(acc2[i], make_float2(fold, fold));
        bias *= fold;
      }
      const float c_eff = __fdividef(sc.y, s_pend);
      const float2 ce2 = make_float2(c_eff, c_eff);
      if (cw == 0) QA_TL(6);
      mbar_wait(&o_full[ob], (j / kOBuf) & 1);
      tc_fence_after();
      if (cw == 0) QA_TL(7);
      if (MG && NDR == 1) {
#pragma unroll
        for (int grp = 0; grp < DCx / 64; ++grp) {                 // two x32 loads in flight per round trip
          uint32_t r[64];
          tmem_ld64(lane_addr + kOCol + ob * 128 + d0 + grp * 64, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {                           // acc += (kMagic + x) * c: one FFMA per element ...
            acc2[grp * 32 + i].x = fmaf(__uint_as_float(r[2 * i]), c_eff, acc2[grp * 32 + i].x);
            acc2[grp * 32 + i].y = fmaf(__uint_as_float(r[2 * i + 1]), c_eff, acc2[grp * 32 + i].y);
          }
        }
        if (!FP8) {
        // ... and the kMagic * c part, identical for every column of the row, is summed on the side and taken out every
        // fourth tile (so that it never outgrows the accumulated values by more than ~2^10: the subtraction then costs
        // < 2^-13 of relative precision, below the fp16 rounding of O)
        bias = fmaf(kMagic, c_eff, bias);
        if ((j & 3) == 3 || j == nk - 1) {
          const float2 nb2 = make_float2(-bias, -bias);
#pragma unroll
          for (int i = 0; i < DCx / 2; ++i) acc2[i] = __fadd2_rn(acc2[i], nb2);
          bias = 0.f;
        }
        }
      } else
#pragma unroll
      for (int ch = 0; ch < DCx / QA_DRAIN_W; ++ch) {
        uint32_t r[QA_DRAIN_W];
        if (QA_DRAIN_W == 32) tmem_ld32(lane_addr + kOCol + ob * 128 + d0 + ch * 32, *reinterpret_cast<uint32_t (*)[32]>(&r[0]));
        else tmem_ld16(lane_addr + kOCol + ob * 128 + d0 + ch * 16, *reinterpret_cast<uint32_t (*)[16]>(&r[0]));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < QA_DRAIN_W / 2; ++i) {
          const float2 x = MG ? __fadd2_rn(make_float2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), make_float2(-kMagic, -kMagic))   // exact
                              : make_float2(__int2float_rn((int)r[2 * i]), __int2float_rn((int)r[2 * i + 1]));
          acc2[ch * (QA_DRAIN_W / 2) + i] = __ffma2_rn(x, ce2, acc2[ch * (QA_DRAIN_W / 2) + i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[ob]);
      if (cw == 0) QA_TL(8);
    }
    mbar_wait(&fin_full, 0);
    float l = l_part[0][row];
    if (NSPLIT == 2) l += l_part[1][row];
    const size_t grow = (size_t)bh * p.Sq + q0 + row;
    if (p.O_acc_out != nullptr) {
      // ring mode: emit the unnormalised accumulator and (m, l); the host-side merge normalises
      float* dst = p.O_acc_out + grow * D + d0;
#pragma unroll
      for (int i = 0; i < DCx; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(acc2[i / 2].x * s_pend, acc2[i / 2].y * s_pend, acc2[i / 2 + 1].x * s_pend, acc2[i / 2 + 1].y * s_pend);
      if (split == 0) { p.m_out[grow] = __half2float(m_fin[row]); p.l_out[grow] = l; }
    } else {
      const float inv_l = s_pend / l;                                          // O / l (:256), pending rescale folded in
      __half* dst = p.O + grow * D + d0;
#pragma unroll
      for (int i = 0; i < DCx; i += 8) {
        __half2 h0 = __floats2half2_rn(acc2[i / 2].x * inv_l, acc2[i / 2].y * inv_l);
        __half2 h1 = __floats2half2_rn(acc2[i / 2 + 1].x * inv_l, acc2[i / 2 + 1].y * inv_l);
        __half2 h2 = __floats2half2_rn(acc2[i / 2 + 2].x * inv_l, acc2[i / 2 + 2].y * inv_l);
        __half2 h3 = __floats2half2_rn(acc2[i / 2 + 3].x * inv_l, acc2[i / 2 + 3].y * inv_l);
        uint4 v;
        v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
        v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + i) = v;
      }
      if (split == 0) {
        const __half m = m_fin[row];
        const float lg = log2f(l);
        p.lse16[grow] = __hadd(m, __float2half_rn(lg));                          // (:252)
        if (p.lse32 != nullptr) p.lse32[grow] = __half2float(m) + lg;
      }
    }
  } else if (warp >= R::kMmaWarp + 1) {
    if (NSPLIT == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");     // idle warps of the last warpgroup
  } else if (warp == R::kTmaWarp) {
    // =========================== TMA producer ===========================
    if (NSPLIT == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (elect_one()) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
      mbar_expect_tx(&q_full, L::kQBytes);
      tma_load_2d(smem + L::off_q, &tm_q, &q_full, 0, bh * p.Sq + q0);
      for (int j = 0; j < nk; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], L::kKBytes);
        tma_load_2d(smem + L::off_k + s * L::kKBytes, &tm_k, &k_full[s], 0, bh * p.Sk + j * kBN);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], L::kVBytes);
        tma_load_2d(smem + L::off_v + s * L::kVBytes, &tm_v, &v_full[s], 0, bh * p.Sk + j * kBN);
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    if (NSPLIT == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (elect_one()) {
      // int8: s32 += s8 x s8; fp8: f32 += e4m3 x e4m3 (format 0).  Q, K K-major; B = V MN-major
      constexpr uint32_t idesc_qk = FP8 ? umma_idesc(1, 0, 0, 0, 0, kBM, kBN) : umma_idesc(2, 1, 1, 0, 0, kBM, kBN);
      constexpr uint32_t idesc_pv = FP8 ? umma_idesc(1, 0, 0, 0, 1, kBM, D) : umma_idesc(2, 1, 1, 0, 1, kBM, D);
      const uint32_t q_addr = smem_u32(smem + L::off_q);
      // accumulator initialisation (MG): D = A B^T over K = 16 fp16 elements of the constant tiles = kMagic everywhere
      constexpr uint32_t idesc_cqk = umma_idesc(1, 0, 0, 0, 0, kBM, kBN), idesc_cpv = umma_idesc(1, 0, 0, 0, 0, kBM, D);
      const uint64_t cdesc_a = umma_smem_desc(smem_u32(smem + L::off_c), 16, 0, kLayoutQK);          // stride 0 between 8-row groups:
      const uint64_t cdesc_b = umma_smem_desc(smem_u32(smem + L::off_c + 1024), 16, 0, kLayoutQK);   // every group reads the same atom
      auto issue_pv = [&](int t) {                                 // Opart[b] = P_t V_t, P from TMEM (S[b] columns)
        const int sb = t % kSBuf, ob = t % kOBuf, s = t % STAGES;
        mbar_wait(&v_full[s], (t / STAGES) & 1);
        mbar_wait(&o_empty[ob], ((t / kOBuf) & 1) ^ 1);
        QA_TLX(t < 64, t, 10);   // V landed, Opart free
        mbar_wait(&p_full[sb], (t / kSBuf) & 1);
        tc_fence_after();
        QA_TLX(t < 64, t, 11);   // PV issue
        const uint32_t v_addr = smem_u32(smem + L::off_v + s * L::kVBytes);
        if (kInit) umma_f16_ss(tbase + kOCol + ob * 128, cdesc_a, cdesc_b, idesc_cpv, 0);  // Opart = kMagic
#pragma unroll
        for (int k = 0; k < kBN / 32; ++k) {
          const uint64_t bd = umma_smem_desc(v_addr + k * 32 * D, 16, kSboQK, kLayoutQK);
          if (FP8) umma_f8_ts(tbase + kOCol + ob * 128, tbase + sb * kSStride + kPOff + k * 8, bd, idesc_pv, k > 0);
          else umma_i8_ts(tbase + kOCol + ob * 128, tbase + sb * kSStride + kPOff + k * 8, bd, idesc_pv, kInit || k > 0);
        }
        umma_commit(&o_full[ob]);
        umma_commit(&v_empty[s]);
      };
      auto issue_qk = [&](int j) {                                 // S[b] = Q K_j^T
        const int sb = j % kSBuf, s = j % STAGES;
        mbar_wait(&k_full[s], (j / STAGES) & 1);
        QA_TLX(j < 64, j, 9);
        const uint32_t k_addr = smem_u32(smem + L::off_k + s * L::kKBytes);
        if (kInit) umma_f16_ss(tbase + sb * kSStride, cdesc_a, cdesc_b, idesc_cqk, 0);          // S = kMagic
#pragma unroll
        for (int k = 0; k < D / 32; ++k) {
          const uint64_t ad = umma_smem_desc(q_addr + k * 32, 16, kSboQK, kLayoutQK);
          const uint64_t bd = umma_smem_desc(k_addr + k * 32, 16, kSboQK, kLayoutQK);
          if (FP8) umma_f8_ss(tbase + sb * kSStride, ad, bd, idesc_qk, k > 0);
          else umma_i8_ss(tbase + sb * kSStride, ad, bd, idesc_qk, kInit || k > 0);
        }
        umma_commit(&s_full[sb]);
        umma_commit(&k_empty[s]);
      };
      mbar_wait(&q_full, 0);
      tc_fence_after();
      for (int j = 0; j < kSBuf && j < nk; ++j) issue_qk(j);
      for (int j = 0; j < nk; ++j) {
        issue_pv(j);
        if (j + kSBuf < nk) issue_qk(j + kSBuf);   // behind P V in the in-order pipe: the S buffer is rewritten only after
      }                                            // its P was consumed
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == R::kMmaWarp) tmem_dealloc<512>(tbase);
}

// Causal row 0 of every head sees no key: it is the uniform average over ALL keys of the de-quantised V (LEDGER B-1:
// what the reference's baseline computes with its finite fill value), lse = -128 + log2(S).
template <int D>
__global__ void __launch_bounds__(512) int8_row0_fixup_kernel(const int8_t* __restrict__ v_i8, const __half* __restrict__ sv,
                                                              __half* O, __half* lse16, float* lse32, int S) {
  constexpr int NSTR = 512 / D;                                 // key stripes
  __shared__ float part[NSTR][D];
  const int bh = blockIdx.x, d = threadIdx.x % D, st = threadIdx.x / D;
  const int ntile = S / 128;
  float acc = 0.f;
  for (int t = st; t < ntile; t += NSTR) {
    const int8_t* vt = v_i8 + ((size_t)bh * S + (size_t)t * 128) * D + d;
    int sum = 0;
#pragma unroll 8
    for (int r = 0; r < 128; ++r) sum += vt[(size_t)r * D];
    acc = fmaf((float)sum, __half2float(sv[(size_t)bh * ntile + t]), acc);
  }
  part[st][d] = acc;
  __syncthreads();
  if (st == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NSTR; ++i) tot += part[i][d];
    O[(size_t)bh * S * D + d] = __float2half_rn(tot / (float)S);
    if (d == 0) {
      const float l = -128.0f + log2f((float)S);
      lse16[(size_t)bh * S] = __float2half_rn(l);
      if (lse32 != nullptr) lse32[(size_t)bh * S] = l;
    }
  }
}

template <int D, int NSPLIT, int STAGES, int BN, bool RN = false, bool CAUSAL = false, bool MG = false, int NDR = 1, bool FP8 = false>
static int launch_int8_fwd(const void* q_i8, const void* k_i8, const void* v_i8, const Int8FwdParams& p, int BH,
                           cudaStream_t st) {
  using L = Int8FwdSmem<D, NSPLIT, STAGES, BN>;
  CUtensorMap tq, tk, tv;
  const int sw = (D == 128) ? 3 : 2;
  uint64_t dq[2] = {(uint64_t)D, (uint64_t)BH * p.Sq}, dk[2] = {(uint64_t)D, (uint64_t)BH * p.Sk};
  uint64_t str[1] = {(uint64_t)D};
  uint32_t box[2] = {(uint32_t)D, 128};
  uint32_t boxk[2] = {(uint32_t)D, (uint32_t)BN};
  int rc;
  if ((rc = qa_make_tmap(&tq, q_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dq, str, box, sw))) return rc;
  if ((rc = qa_make_tmap(&tk, k_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dk, str, boxk, sw))) return rc;
  if ((rc = qa_make_tmap(&tv, v_i8, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dk, str, boxk, sw))) return rc;
  auto kern = int8_fwd_kernel<D, NSPLIT, STAGES, BN, RN, CAUSAL, MG, NDR, FP8>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  dim3 grid(p.Sq / kBM, BH);
  kern<<<grid, Int8FwdRoles<NSPLIT, NDR>::kThreads, L::total, st>>>(tq, tk, tv, p);
  return qa_check_launch("qa_int8_fwd");
}

}  // namespace qa

using namespace qa;

#ifdef QA_DEV_TIMELINE
static void* g_int8_fwd_dbg = nullptr;
// Development library only (include/qattn_dev.h): CTA (0,0) of subsequent qa_int8_fwd launches records SM-clock stamps
// per k-tile into buf ([64 tiles][16 slots] int64); NULL switches it off.  Not thread-safe; used by tools/timeline.py.
extern "C" int qa_debug_set_int8_fwd_timeline(void* buf) {
  g_int8_fwd_dbg = buf;
  return 0;
}
#endif

// Forward over pre-quantised operands.  q_i8 [BH*Sq, D], k_i8 / v_i8 [BH*Sk, D] int8 row-major; sq [BH*Sq/Bq],
// sk / sv [BH*Sk/Bkv] fp16.  Outputs: O fp16 [BH*Sq, D], lse16 fp16 [BH*Sq], lse32 fp32 [BH*Sq] (optional).
// Ring mode (o_acc != NULL): writes unnormalised fp32 O plus (m, l) per row instead of O / lse.
extern "C" int qa_int8_fwd_ragged(const void* q_i8, const void* k_i8, const void* v_i8, const void* sq, const void* sk,
                                  const void* sv, void* O, void* lse16, void* lse32, void* o_acc, void* m_out, void* l_out,
                                  const void* o_acc_in, const void* m_in, const void* l_in, int BH, int Sq, int Sk, int Sk_valid,
                                  int D, int Bq, int Bkv, int nsplit, int flags, void* stream) {
  if (Sk_valid <= 0 || Sk_valid > Sk) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: Sk_valid must be in (0, Sk]");
  if (Sk - Sk_valid >= (Bkv > 128 ? Bkv : 128)) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: pad the sequence to the NEXT multiple of max(128, Bkv) only");
  if (Sk_valid != Sk && (flags & QA_FLAG_CAUSAL)) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: ragged sequences are built for the non-causal kernels");
  if (flags & ~(QA_FLAG_NEAREST | QA_FLAG_CAUSAL)) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: unknown flag bits");
  const int rounding = (flags & QA_FLAG_NEAREST) ? 1 : 0;
  const bool causal = (flags & QA_FLAG_CAUSAL) != 0;
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: D must be 64 or 128");
  if (Bkv != 32 && Bkv != 64 && Bkv != 128 && Bkv != 256) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: Bkv must be 32, 64, 128 or 256");
  if (Bq != 32 && Bq != 64 && Bq != 128 && Bq != 256) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: Bq must be 32/64/128/256");
  if (Sq % 128 || Sk % Bkv || Sq % Bq) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: Sq must be a multiple of 128 (and of Bq), Sk of Bkv");
  if (((uintptr_t)q_i8 | (uintptr_t)k_i8 | (uintptr_t)v_i8 | (uintptr_t)O | (uintptr_t)o_acc) & 15)
    return qa_fail(QA_ERR_ALIGN, "qa_int8_fwd: 16-byte alignment required");
  Int8FwdParams p;
  p.sq = (const __half*)sq; p.sk = (const __half*)sk; p.sv = (const __half*)sv;
  p.O = (__half*)O; p.lse16 = (__half*)lse16; p.lse32 = (float*)lse32;
  p.O_acc_out = (float*)o_acc; p.m_out = (float*)m_out; p.l_out = (float*)l_out;
  p.O_acc_in = (const float*)o_acc_in; p.m_in = (const float*)m_in; p.l_in = (const float*)l_in;
  p.Sq = Sq; p.Sk = Sk; p.Bq = Bq; p.Sk_valid = Sk_valid;
  p.qk_scale = (float)((1.0 / sqrt((double)D)) * 1.44269504);
#ifdef QA_DEV_TIMELINE
  p.dbg = (long long*)g_int8_fwd_dbg;
#else
  p.dbg = nullptr;
#endif
  cudaStream_t st = (cudaStream_t)stream;
  // nsplit: 0 = default kernel for the tile; 1 = single-stage softmax (one thread per row); 2 = two-stage softmax without
  // magic accumulators (A/B comparison).  Bkv = 128 defaults to the two-stage kernel with magic accumulators.
  if (nsplit < 0 || nsplit > 3) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: nsplit must be 0, 1, 2 or 3");
  if (causal) {                                                  // instantiated for the tuned tile only
    if (Bkv != 128 || Bq != 128 || nsplit == 1 || rounding || Sq != Sk || o_acc_in != nullptr)
      return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: causal needs Sq == Sk, Bq = Bkv = 128, the two-stage kernel, truncation, no incoming state");
    int rc = D == 128 ? launch_int8_fwd<128, 2, 3, 128, false, true, true>(q_i8, k_i8, v_i8, p, BH, st)
                      : launch_int8_fwd<64, 2, 4, 128, false, true, true>(q_i8, k_i8, v_i8, p, BH, st);
    if (rc) return rc;
    // state-out mode (o_acc != NULL; the diagonal chunk of a causal ring): a row without a visible key leaves the fresh state
    // (O = 0, m = -inf, l = 1) and the caller owns the row-0 rule, which needs the V of EVERY rank
    if (o_acc != nullptr) return QA_OK;
    if (D == 128) int8_row0_fixup_kernel<128><<<BH, 512, 0, st>>>((const int8_t*)v_i8, p.sv, p.O, p.lse16, p.lse32, Sq);
    else int8_row0_fixup_kernel<64><<<BH, 512, 0, st>>>((const int8_t*)v_i8, p.sv, p.O, p.lse16, p.lse32, Sq);
    return qa_check_launch("qa_int8_fwd(causal row 0)");
  }
  if (Bkv == 128) {
    if (rounding == 1) {                                         // accuracy mode: instantiated for the tuned tile only
      if (nsplit == 1) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: nearest rounding needs the two-stage kernel");
      return D == 128 ? launch_int8_fwd<128, 2, 3, 128, true, false, true>(q_i8, k_i8, v_i8, p, BH, st)
                      : launch_int8_fwd<64, 2, 4, 128, true, false, true>(q_i8, k_i8, v_i8, p, BH, st);
    }
    if (nsplit == 3)                                             // same with two drain warps per 32-row group (24 warps)
      return D == 128 ? launch_int8_fwd<128, 2, 3, 128, false, false, true, 2>(q_i8, k_i8, v_i8, p, BH, st)
                      : launch_int8_fwd<64, 2, 4, 128, false, false, true, 2>(q_i8, k_i8, v_i8, p, BH, st);
    if (nsplit == 0)                                             // two-stage kernel with magic accumulators
      return D == 128 ? launch_int8_fwd<128, 2, 3, 128, false, false, true>(q_i8, k_i8, v_i8, p, BH, st)
                      : launch_int8_fwd<64, 2, 4, 128, false, false, true>(q_i8, k_i8, v_i8, p, BH, st);
    if (D == 128) return nsplit == 2 ? launch_int8_fwd<128, 2, 3, 128>(q_i8, k_i8, v_i8, p, BH, st)
                                     : launch_int8_fwd<128, 1, 3, 128>(q_i8, k_i8, v_i8, p, BH, st);
    return nsplit == 2 ? launch_int8_fwd<64, 2, 4, 128>(q_i8, k_i8, v_i8, p, BH, st)
                       : launch_int8_fwd<64, 1, 4, 128>(q_i8, k_i8, v_i8, p, BH, st);
  }
  if (rounding == 1) return qa_fail(QA_ERR_SHAPE, "qa_int8_fwd: nearest rounding needs Bkv = 128");
  if (Bkv == 256) return D == 128 ? launch_int8_fwd<128, 1, 2, 256>(q_i8, k_i8, v_i8, p, BH, st)
                                  : launch_int8_fwd<64, 1, 2, 256>(q_i8, k_i8, v_i8, p, BH, st);
  if (Bkv == 64) return D == 128 ? launch_int8_fwd<128, 1, 4, 64>(q_i8, k_i8, v_i8, p, BH, st)
                                 : launch_int8_fwd<64, 1, 4, 64>(q_i8, k_i8, v_i8, p, BH, st);
  return D == 128 ? launch_int8_fwd<128, 1, 4, 32>(q_i8, k_i8, v_i8, p, BH, st)
                  : launch_int8_fwd<64, 1, 4, 32>(q_i8, k_i8, v_i8, p, BH, st);
}

extern "C" int qa_int8_fwd_state(const void* q_i8, const void* k_i8, const void* v_i8, const void* sq, const void* sk,
                                 const void* sv, void* O, void* lse16, void* lse32, void* o_acc, void* m_out, void* l_out,
                                 const void* o_acc_in, const void* m_in, const void* l_in, int BH, int Sq, int Sk, int D,
                                 int Bq, int Bkv, int nsplit, int flags, void* stream) {
  return qa_int8_fwd_ragged(q_i8, k_i8, v_i8, sq, sk, sv, O, lse16, lse32, o_acc, m_out, l_out, o_acc_in, m_in, l_in, BH, Sq, Sk,
                            Sk, D, Bq, Bkv, nsplit, flags, stream);
}

extern "C" int qa_int8_fwd(const void* q_i8, const void* k_i8, const void* v_i8, const void* sq, const void* sk,
                           const void* sv, void* O, void* lse16, void* lse32, void* o_acc, void* m_out, void* l_out,
                           int BH, int Sq, int Sk, int D, int Bq, int Bkv, int nsplit, int flags, void* stream) {
  return qa_int8_fwd_ragged(q_i8, k_i8, v_i8, sq, sk, sv, O, lse16, lse32, o_acc, m_out, l_out, nullptr, nullptr, nullptr, BH,
                            Sq, Sk, Sk, D, Bq, Bkv, nsplit, flags, stream);
}

// fp8 (e4m3) forward over pre-quantised operands: q / k / v bytes are e4m3 produced by qa_quant_block with
// rounding = 2 (scale = amax / 448), Bq = Bkv = 128.  Same outputs as qa_int8_fwd (normal mode).  Forward only.
extern "C" int qa_fp8_fwd(const void* q_e4m3, const void* k_e4m3, const void* v_e4m3, const void* sq, const void* sk,
                          const void* sv, void* O, void* lse16, void* lse32, int BH, int Sq, int Sk, int D, void* stream) {
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_fp8_fwd: D must be 64 or 128");
  if (BH <= 0 || Sq <= 0 || Sk <= 0 || Sq % 128 || Sk % 128) return qa_fail(QA_ERR_SHAPE, "qa_fp8_fwd: Sq, Sk must be positive multiples of 128");
  if (!q_e4m3 || !k_e4m3 || !v_e4m3 || !sq || !sk || !sv || !O || !lse16) return qa_fail(QA_ERR_SHAPE, "qa_fp8_fwd: null pointer");
  if (((uintptr_t)q_e4m3 | (uintptr_t)k_e4m3 | (uintptr_t)v_e4m3 | (uintptr_t)O) & 15) return qa_fail(QA_ERR_ALIGN, "qa_fp8_fwd: 16-byte alignment required");
  Int8FwdParams p;
  p.sq = (const __half*)sq; p.sk = (const __half*)sk; p.sv = (const __half*)sv;
  p.O = (__half*)O; p.lse16 = (__half*)lse16; p.lse32 = (float*)lse32;
  p.O_acc_out = nullptr; p.m_out = nullptr; p.l_out = nullptr; p.O_acc_in = nullptr; p.m_in = nullptr; p.l_in = nullptr;
  p.Sq = Sq; p.Sk = Sk; p.Bq = 128; p.Sk_valid = Sk;
  p.qk_scale = (float)((1.0 / sqrt((double)D)) * 1.44269504);
  p.dbg = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  return D == 128 ? launch_int8_fwd<128, 2, 3, 128, true, false, true, 1, true>(q_e4m3, k_e4m3, v_e4m3, p, BH, st)
                  : launch_int8_fwd<64, 2, 4, 128, true, false, true, 1, true>(q_e4m3, k_e4m3, v_e4m3, p, BH, st);
}
