// bf16 flash attention forward with the Qiu & Yao bias-corrected running max (reference
// attention_bf16.py:195-294) for sm_100a.  Warp roles: softmax warps, one correction warpgroup, TMA producer warp,
// MMA warp.  S = Q K^T : tcgen05 kind::f16, fp16 x fp16 -> fp32 TMEM (double-buffered).  O += P V : bf16 x bf16,
// the fp32 accumulator stays RESIDENT in TMEM across k-tiles; the correction warps multiply it by the bf16 rescale
// factor only when a row's running max moved (rescale == 1 exactly otherwise), so the steady-state TMEM->register
// traffic is the S tile alone (TMEM read bandwidth, ~100 B/clk/SM measured, is what bounds these kernels).
// Per k-tile numerics (contract mode): u = bf16(S * qk_scale) (one rounding of the fp32 logit); strict causal mask with
// masked weight 0; m' = max(m, rowmax(u)) in bf16; if >= 2 entries lie within 1e-3 of m' (one scaled domain):
// m' = beta*m' (m' > 0) or 0 (m' < 0); P = bf16(exp2(bf16(u - m'))); l = l*rescale + sum(P);
// rescale = bf16(exp2(bf16(m - m'))); O = O*rescale + P V.  Output O fp32, lse = m + log2(l) fp32.
#include "qa_ptx.cuh"
#include "qa_host.h"
#include <type_traits>

namespace qa {

constexpr int kAtom = 128 * 128;           // one 128-row x 128-byte swizzle atom column: 16 KiB

template <int D, int STAGES, int PBUF>
struct Bf16FwdSmem {
  static constexpr int kTile = 128 * D * 2;        // Q / K / V tile bytes (16-bit)
  static constexpr int kPBytes = 128 * 128 * 2;
  static constexpr int off_q = 0;
  static constexpr int off_k = off_q + kTile;
  static constexpr int off_v = off_k + STAGES * kTile;
  static constexpr int off_p = off_v + STAGES * kTile;
  static constexpr int total = off_p + PBUF * kPBytes + 1024;
};

struct Bf16FwdParams {
  float* O;        // [BH*Sq, D] fp32
  float* lse;      // [BH*Sq] fp32
  int Sq, Sk, causal, BH;
  int Sk_valid;        // keys [Sk_valid, Sk) of every head are zero padding (ragged sequence): weight exactly 0
  float qk_scale;
  float rescale_tau;   // see qa_bf16_fwd_ex
};

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {      // RN, a -> low half
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ __nv_bfloat162 as_bf2(uint32_t v) { return *reinterpret_cast<__nv_bfloat162*>(&v); }
__device__ __forceinline__ uint32_t as_u32(__nv_bfloat162 v) { return *reinterpret_cast<uint32_t*>(&v); }

template <int D, int NSPLIT, int STAGES, int PBUF>
__global__ void __launch_bounds__(128 * NSPLIT + 192, 1)
bf16_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, Bf16FwdParams p) {
  using L = Bf16FwdSmem<D, STAGES, PBUF>;
  constexpr int NC = 128 / NSPLIT;
  constexpr int kSoftWarps = 4 * NSPLIT;
  constexpr int kProdWarp = kSoftWarps + 4;
  constexpr int kMmaWarp = kSoftWarps + 5;
  constexpr int kDAtoms = D / 64;              // 128-byte atoms along D for 16-bit operands

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t q_full, k_full[STAGES], k_empty[STAGES], v_full[STAGES], v_empty[STAGES];
  __shared__ uint64_t s_full[2], s_empty[2], p_full[2], p_empty[2], o_full, o_ready, sc_full[2], sc_empty[2], fin_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ float row_sc[2][128];            // rescale per row
  __shared__ uint32_t xtop[2][2][128];        // NSPLIT == 2: packed (top1, top2) bf16 exchange
  __shared__ float l_part[2][128];
  __shared__ float m_fin[128];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nqt = p.Sq / 128;
  int rank, bh;
  qa_group_order((int)blockIdx.x, p.BH, nqt, p.causal ? 16 : 1, rank, bh);
  const int qt = nqt - 1 - rank;                               // heaviest (latest) causal tiles first
  const int q0 = qt * 128;
  const int nkv = (p.Sk_valid + 127) / 128;                    // k-tiles without a valid key are skipped
  const int nk = p.causal ? min(nkv, qt + 1) : nkv;

  if (tid == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], kSoftWarps);
      mbar_init(&p_full[b], kSoftWarps); mbar_init(&p_empty[b], 1);
      mbar_init(&sc_full[b], 4); mbar_init(&sc_empty[b], 4);
    }
    mbar_init(&o_full, 1); mbar_init(&o_ready, 4);
    mbar_init(&fin_full, kSoftWarps);
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp < kSoftWarps) {
    // =========================== softmax warps ===========================
    const int split = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const int c0 = split * NC;
    const int grow = q0 + row;                                    // query index inside the head
    __nv_bfloat16 m_bf = __float2bfloat16(-INFINITY);
    float l = (split == 0) ? 1.0f : 0.0f;                        // attention_bf16.py:198
    const uint32_t ninf2 = 0xff80ff80u;                          // (-inf, -inf) bf16x2
    for (int j = 0; j < nk; ++j) {
      const int b = j & 1;
      const int pb = (PBUF == 2) ? b : 0;
      const uint32_t ph = (j >> 1) & 1;
      const uint32_t pph = (PBUF == 2) ? ph : (j & 1);
      const bool diag = p.causal && (j == qt);
      const bool tail = (j + 1) * 128 > p.Sk_valid;                // ragged last k-tile
      const int klim = diag ? min(grow, p.Sk_valid) : p.Sk_valid;  // first key without weight
      mbar_wait(&s_full[b], ph);
      tc_fence_after();
      // ---- pass 1: u = bf16(bf16(S) * qk_scale), masked; per-thread top-2
      uint32_t u2[NC / 2];
      __nv_bfloat162 t1 = as_bf2(ninf2), t2 = as_bf2(ninf2);
      auto pass1 = [&](auto masked) {                              // two instantiations: the mask costs nothing off-diagonal
#pragma unroll
        for (int ch = 0; ch < NC / 32; ++ch) {
          uint32_t r[32];
          tmem_ld32(lane_addr + b * 128 + c0 + ch * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            uint32_t u = pack_bf16(__uint_as_float(r[2 * i]) * p.qk_scale, __uint_as_float(r[2 * i + 1]) * p.qk_scale);   // u = bf16(S * qk_scale)
            if (decltype(masked)::value) {                         // strict causal: keep key < query; ragged: keep key < Sk_valid
              const int key = j * 128 + c0 + ch * 32 + 2 * i;
              if (key >= klim) u = (u & 0xffff0000u) | 0xff80u;
              if (key + 1 >= klim) u = (u & 0x0000ffffu) | 0xff800000u;
            }
            u2[ch * 16 + i] = u;
            const __nv_bfloat162 x = as_bf2(u);
            t2 = __hmax2(t2, __hmin2(t1, x));
            t1 = __hmax2(t1, x);
          }
        }
      };
      if (diag || tail) pass1(std::true_type{}); else pass1(std::false_type{});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[b]);
      // combine the two lanes of the packed top-2 trackers
      __nv_bfloat16 a1 = __low2bfloat16(t1), b1 = __high2bfloat16(t1), a2 = __low2bfloat16(t2), b2 = __high2bfloat16(t2);
      __nv_bfloat16 top1 = __hmax(a1, b1);
      __nv_bfloat16 top2 = __hmax(__hmin(a1, b1), __hmax(a2, b2));
      if (NSPLIT == 2) {
        xtop[b][split][row] = (uint32_t)__bfloat16_as_ushort(top1) | ((uint32_t)__bfloat16_as_ushort(top2) << 16);
        named_bar_sync(1 + (warp & 3), 32 * NSPLIT);   // only the warps that share these 32 rows meet
        const uint32_t o = xtop[b][split ^ 1][row];
        const __nv_bfloat16 o1 = __ushort_as_bfloat16((unsigned short)(o & 0xffff)), o2 = __ushort_as_bfloat16((unsigned short)(o >> 16));
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
      if (!(__bfloat162float(__hsub(m_new, m_bf)) > p.rescale_tau)) m_new = m_bf;     // lazy rescale (qa_bf16_fwd_ex)
      const float resc = __bfloat162float(__float2bfloat16(ex2_approx(__bfloat162float(__hsub(m_bf, m_new)))));
      m_bf = m_new;
      if (split == 0 && j > 0) {                                 // tile 0 overwrites O: no rescale to hand over
        const int sb = (j - 1) & 1;
        mbar_wait(&sc_empty[sb], (((j - 1) >> 1) & 1) ^ 1);
        row_sc[sb][row] = resc;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sc_full[sb]);
      }
      // ---- pass 2: P = bf16(exp2(bf16(u - m'))), l += sum(P), P -> smem (K-major, two 64-key atoms)
      mbar_wait(&p_empty[pb], pph ^ 1);
      const __nv_bfloat162 m2 = __bfloat162bfloat162(m_new);
      float lsum = 0.f;
      uint8_t* pbase = smem + L::off_p + pb * L::kPBytes;
#pragma unroll
      for (int g = 0; g < NC / 8; ++g) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t t = as_u32(__hsub2(as_bf2(u2[g * 4 + e]), m2));
          const uint32_t pp = pack_bf16(ex2_approx(bf_lo(t)), ex2_approx(bf_hi(t)));
          lsum += bf_lo(pp) + bf_hi(pp);
          w[e] = pp;
        }
        const int col = c0 + g * 8;
        const uint32_t off = (uint32_t)(col >> 6) * kAtom + swz128(row, (col & 63) * 2);
        sts128(smem_u32(pbase) + off, w[0], w[1], w[2], w[3]);
      }
      l = l * resc + lsum;
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[pb]);
    }
    l_part[split][row] = l;
    if (split == 0) m_fin[row] = __bfloat162float(m_bf);
    __syncwarp();
    if (lane == 0) mbar_arrive(&fin_full);
  } else if (warp < kProdWarp) {
    // =========================== correction warpgroup (4 warps, thread = row) ===========================
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    for (int j = 1; j < nk; ++j) {                               // tile 0 overwrites O (accumulate = 0): nothing to rescale
      const int sb = (j - 1) & 1;
      mbar_wait(&sc_full[sb], ((j - 1) >> 1) & 1);
      const float resc = row_sc[sb][row];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sc_empty[sb]);
      if (__any_sync(0xffffffffu, resc != 1.0f)) {               // some row's running max moved: O *= rescale (:280)
        mbar_wait(&o_full, (j - 1) & 1);                         // P V of tile j-1 has landed in TMEM
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < D / 32; ++ch) {
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
      if (lane == 0) mbar_arrive(&o_ready);                      // the MMA warp may accumulate tile j
    }
    mbar_wait(&fin_full, 0);
    mbar_wait(&o_full, (nk - 1) & 1);
    tc_fence_after();
    float l = l_part[0][row];
    if (NSPLIT == 2) l += l_part[1][row];
    const size_t gr = (size_t)bh * p.Sq + q0 + row;
    const float inv_l = 1.0f / l;
    float* dst = p.O + gr * D;
#pragma unroll
    for (int ch = 0; ch < D / 32; ++ch) {
      uint32_t r[32];
      tmem_ld32(lane_addr + 256 + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 4)
        *reinterpret_cast<float4*>(dst + ch * 32 + i) =
            make_float4(__uint_as_float(r[i]) * inv_l, __uint_as_float(r[i + 1]) * inv_l, __uint_as_float(r[i + 2]) * inv_l,
                        __uint_as_float(r[i + 3]) * inv_l);
    }
    p.lse[gr] = m_fin[row] + log2f(l);                                       // attention_bf16.py:288
  } else if (warp == kProdWarp) {
    // =========================== TMA producer ===========================
    if (elect_one()) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
      mbar_expect_tx(&q_full, L::kTile);
#pragma unroll
      for (int a = 0; a < kDAtoms; ++a) tma_load_2d(smem + L::off_q + a * kAtom, &tm_q, &q_full, a * 64, bh * p.Sq + q0);
      for (int j = 0; j < nk; ++j) {
        const int s = j % STAGES;
        const uint32_t ph = (j / STAGES) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], L::kTile);
#pragma unroll
        for (int a = 0; a < kDAtoms; ++a)
          tma_load_2d(smem + L::off_k + s * L::kTile + a * kAtom, &tm_k, &k_full[s], a * 64, bh * p.Sk + j * 128);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], L::kTile);
#pragma unroll
        for (int a = 0; a < kDAtoms; ++a)
          tma_load_2d(smem + L::off_v + s * L::kTile + a * kAtom, &tm_v, &v_full[s], a * 64, bh * p.Sk + j * 128);
      }
    }
  } else {
    // =========================== MMA issuer ===========================
    if (elect_one()) {
      constexpr uint32_t idesc_qk = umma_idesc(1, 0, 0, 0, 0, 128, 128);        // f32 += f16 x f16, K-major
      constexpr uint32_t idesc_pv = umma_idesc(1, 1, 1, 0, 1, 128, D);          // f32 += bf16 x bf16, B = V MN-major
      const uint32_t q_addr = smem_u32(smem + L::off_q);
      auto issue_pv = [&](int t) {
        const int b = t & 1, s = t % STAGES;
        const int pb = (PBUF == 2) ? b : 0;
        const uint32_t ph = (t >> 1) & 1;
        const uint32_t pph = (PBUF == 2) ? ph : (t & 1);
        mbar_wait(&v_full[s], (t / STAGES) & 1);
        if (t > 0) mbar_wait(&o_ready, (t - 1) & 1);                              // O rescaled (if needed) for tile t
        mbar_wait(&p_full[pb], pph);
        tc_fence_after();
        const uint32_t p_addr = smem_u32(smem + L::off_p + pb * L::kPBytes);
        const uint32_t v_addr = smem_u32(smem + L::off_v + s * L::kTile);
#pragma unroll
        for (int k = 0; k < 8; ++k) {                                            // 128 keys / 16
          const uint64_t ad = umma_smem_desc(p_addr + (k >> 2) * kAtom + (k & 3) * 32, 16, 1024, kSwz128);
          const uint64_t bd = umma_smem_desc(v_addr + k * 2048, kAtom, 1024, kSwz128);
          umma_f16_ss(tbase + 256, ad, bd, idesc_pv, (t > 0) || (k > 0));       // O stays resident in TMEM
        }
        umma_commit(&o_full);
        umma_commit(&v_empty[s]);
        umma_commit(&p_empty[pb]);
      };
      mbar_wait(&q_full, 0);
      for (int j = 0; j < nk; ++j) {
        const int b = j & 1, s = j % STAGES;
        mbar_wait(&k_full[s], (j / STAGES) & 1);
        mbar_wait(&s_empty[b], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + L::off_k + s * L::kTile);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t o = (k >> 2) * kAtom + (k & 3) * 32;
          umma_f16_ss(tbase + b * 128, umma_smem_desc(q_addr + o, 16, 1024, kSwz128), umma_smem_desc(k_addr + o, 16, 1024, kSwz128),
                      idesc_qk, k > 0);
        }
        umma_commit(&s_full[b]);
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

// causal row 0 has no visible key: reference kernel and baseline both produce the uniform average over ALL keys
// (LEDGER B-1).  grid = BH, block = 1024 threads: D/8 threads across the head dim (16-byte loads), the rest across keys.
template <int D>
__global__ void __launch_bounds__(1024) bf16_row0_fixup_kernel(const __nv_bfloat16* __restrict__ v, float* __restrict__ O,
                                                              float* __restrict__ lse, int Sq, int Sk, int Sk_valid) {
  constexpr int TX = D / 8, TY = 1024 / TX;
  __shared__ float red[TY][D + 1];
  const int bh = blockIdx.x, tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const uint4* base = reinterpret_cast<const uint4*>(v + (size_t)bh * Sk * D);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = ty; k < Sk_valid; k += TY) {
    const uint4 t = __ldg(base + (size_t)k * TX + tx);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[2 * i] += bf_lo(w[i]); acc[2 * i + 1] += bf_hi(w[i]); }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < D) {
    float s = 0.f;
    for (int r = 0; r < TY; ++r) s += red[r][threadIdx.x];
    O[(size_t)bh * Sq * D + threadIdx.x] = s / (float)Sk_valid;
    if (threadIdx.x == 0) lse[(size_t)bh * Sq] = -128.0f + log2f((float)Sk_valid);
  }
}

static int launch_row0_fixup(const void* v, float* O, float* lse, int BH, int Sq, int Sk, int Sk_valid, int D, cudaStream_t st) {
  if (D == 128) bf16_row0_fixup_kernel<128><<<BH, 1024, 0, st>>>((const __nv_bfloat16*)v, O, lse, Sq, Sk, Sk_valid);
  else bf16_row0_fixup_kernel<64><<<BH, 1024, 0, st>>>((const __nv_bfloat16*)v, O, lse, Sq, Sk, Sk_valid);
  return qa_check_launch("qa_bf16_fwd(row0)");
}

template <int D, int NSPLIT, int STAGES, int PBUF>
static int launch_bf16_fwd(const void* q, const void* k, const void* v, const Bf16FwdParams& p, int BH, cudaStream_t st) {
  using L = Bf16FwdSmem<D, STAGES, PBUF>;
  CUtensorMap tq, tk, tv;
  uint64_t dq[2] = {(uint64_t)D, (uint64_t)BH * p.Sq}, dk[2] = {(uint64_t)D, (uint64_t)BH * p.Sk};
  uint64_t str[1] = {(uint64_t)D * 2};
  uint32_t box[2] = {64, 128};
  int rc;
  if ((rc = qa_make_tmap(&tq, q, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dq, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tk, k, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dk, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tv, v, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dk, str, box, 3))) return rc;
  auto kern = bf16_fwd_kernel<D, NSPLIT, STAGES, PBUF>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  dim3 grid((p.Sq / 128) * BH);                       // order: qa_group_order
  kern<<<grid, 128 * NSPLIT + 192, L::total, st>>>(tq, tk, tv, p);
  int r = qa_check_launch("qa_bf16_fwd");
  if (r) return r;
  if (p.causal) r = launch_row0_fixup(v, p.O, p.lse, BH, p.Sq, p.Sk, p.Sk_valid, D, st);
  return r;
}

// two-query-tile variant (attn_bf16_fwd2.cu)
template <int D, int STAGES>
int launch_bf16_fwd2(const void* q, const void* k, const void* v, float* O, float* lse, int BH, int Sq, int Sk, int Sk_valid,
                     int causal, float qk_scale, float rescale_tau, cudaStream_t st);

}  // namespace qa

using namespace qa;

// q, k: fp16 [BH*S, D]; v: bf16 [BH*Sk, D]; O: fp32 [BH*Sq, D]; lse: fp32 [BH*Sq] (log2-sum-exp2).
// Ragged sequences (the reference's hl.tile clamps the last tile, attention_bf16.py:170,201): the buffers are zero-padded
// per head to Sq / Sk (multiples of 128); keys [Sk_valid, Sk) have weight exactly 0.
extern "C" int qa_bf16_fwd_ragged(const void* q_f16, const void* k_f16, const void* v_bf16, void* O_f32, void* lse_f32, int BH,
                                  int Sq, int Sk, int Sk_valid, int D, int causal, int nsplit, float rescale_tau, void* stream) {
  if (!(rescale_tau >= 0.f && rescale_tau <= 16.f)) return qa_fail(QA_ERR_SHAPE, "qa_bf16_fwd: rescale_tau must be in [0, 16]");
  if (D != 64 && D != 128) return qa_fail(QA_ERR_SHAPE, "qa_bf16_fwd: D must be 64 or 128");
  if (Sq % 128 || Sk % 128 || Sq <= 0 || Sk <= 0 || BH <= 0) return qa_fail(QA_ERR_SHAPE, "qa_bf16_fwd: Sq, Sk must be positive multiples of 128");
  if (Sk_valid <= Sk - 128 || Sk_valid > Sk) return qa_fail(QA_ERR_SHAPE, "qa_bf16_fwd: Sk_valid must lie in (Sk - 128, Sk]");
  if (causal && Sq != Sk) return qa_fail(QA_ERR_SHAPE, "qa_bf16_fwd: causal needs Sq == Sk");
  if (!q_f16 || !k_f16 || !v_bf16 || !O_f32 || !lse_f32) return qa_fail(QA_ERR_ALIGN, "qa_bf16_fwd: null pointer");
  if (((uintptr_t)q_f16 | (uintptr_t)k_f16 | (uintptr_t)v_bf16 | (uintptr_t)O_f32) & 15)
    return qa_fail(QA_ERR_ALIGN, "qa_bf16_fwd: 16-byte alignment required");
  Bf16FwdParams p;
  p.O = (float*)O_f32; p.lse = (float*)lse_f32; p.Sq = Sq; p.Sk = Sk; p.Sk_valid = Sk_valid; p.causal = causal; p.BH = BH; p.rescale_tau = rescale_tau;
  p.qk_scale = (float)((1.0 / sqrt((double)D)) * 1.44269504);
  cudaStream_t st = (cudaStream_t)stream;
  if ((nsplit == 0 || nsplit == 3) && Sq % 256 == 0) {           // default schedule: two query tiles per CTA, P and O in TMEM
    int rc = D == 128 ? launch_bf16_fwd2<128, 2>(q_f16, k_f16, v_bf16, p.O, p.lse, BH, Sq, Sk, Sk_valid, causal, p.qk_scale, rescale_tau, st)
                      : launch_bf16_fwd2<64, 3>(q_f16, k_f16, v_bf16, p.O, p.lse, BH, Sq, Sk, Sk_valid, causal, p.qk_scale, rescale_tau, st);
    if (rc) return rc;
    if (causal) rc = launch_row0_fixup(v_bf16, p.O, p.lse, BH, Sq, Sk, Sk_valid, D, st);
    return rc;
  }
  if (nsplit == 0 || nsplit == 3) nsplit = 2;
  if (D == 128) return nsplit == 2 ? launch_bf16_fwd<128, 2, 2, 1>(q_f16, k_f16, v_bf16, p, BH, st)
                                   : launch_bf16_fwd<128, 1, 2, 1>(q_f16, k_f16, v_bf16, p, BH, st);
  return nsplit == 2 ? launch_bf16_fwd<64, 2, 3, 2>(q_f16, k_f16, v_bf16, p, BH, st)
                     : launch_bf16_fwd<64, 1, 3, 2>(q_f16, k_f16, v_bf16, p, BH, st);
}

extern "C" int qa_bf16_fwd_ex(const void* q_f16, const void* k_f16, const void* v_bf16, void* O_f32, void* lse_f32, int BH,
                              int Sq, int Sk, int D, int causal, int nsplit, float rescale_tau, void* stream) {
  return qa_bf16_fwd_ragged(q_f16, k_f16, v_bf16, O_f32, lse_f32, BH, Sq, Sk, Sk, D, causal, nsplit, rescale_tau, stream);
}

// Default schedule and the default lazy-rescale threshold (8 log2 units, i.e. P <= 256).
extern "C" int qa_bf16_fwd(const void* q_f16, const void* k_f16, const void* v_bf16, void* O_f32, void* lse_f32, int BH,
                           int Sq, int Sk, int D, int causal, int nsplit, void* stream) {
  return qa_bf16_fwd_ex(q_f16, k_f16, v_bf16, O_f32, lse_f32, BH, Sq, Sk, D, causal, nsplit, 8.0f, stream);
}
