// Warp-specialised recompute backward of the bf16 path for D = 128 (reference
// attention_bf16.py:361-444 under the 8-LEDGER contract B-4..B-8).  Same arithmetic as bf16_bwd_kernel
// (attn_bf16_bwd.cu): P rounded to bf16 for dV, dS = P*(dP - delta) in fp16 (P and dP - delta rounded to fp16, product
// rounded to fp16: 3 x 2^-12 relative where the sequential kernel has 2^-12), fp32 accumulation.
//
// Work item = one 128-key tile j of one (batch, head), looping over the query tiles i (k-outer, as the reference); a
// CTA that has finished its item takes over the item of a CTA that has not started yet (cluster launch control), so the
// dV / dK write-back of one item and the loads of the next overlap.  The logits are computed TRANSPOSED so that the probabilities can feed the tensor core from TMEM:
//   S^T  = K_j Q_i^T        (fp16)  -> TMEM A   [lane = key, column = query]
//   dP^T = V_j dO_i^T       (bf16)  -> TMEM B
//   dV_j += P^T dO_i        (bf16)  P^T read from TMEM (written over S^T by the compute warps)  -> TMEM, resident
//   dK_j += dS^T Q_i        (fp16)  dS^T from shared memory (K-major A)                          -> TMEM, resident
//   dQ_i  = dS K_j          (fp16)  dS^T as MN-major A, K as MN-major B -> TMEM B [lane = query, column = d], staged
//                                   through shared memory and added to global dQ by TMA reduce-add
// Shared memory (7 x 32 KB): K, V, two stages of Q, one of dO (its life ends with dV, early in the tile), dS^T, the dQ
// staging tile (dQ reduce-adds and the dV / dK stores of a finished item).  P never touches shared memory.
// Roles (16 warps, setmaxnreg): warps 0..3 drain dQ (and dV / dK at the end of an item), 4..11 compute P^T and dS^T (two per TMEM lane quadrant, 64 query
// columns each), 12 issues tcgen05.mma, 13 issues the TMA loads.  Issue order of the MMA warp per query tile n:
//   dV(n) | S(n+1) | dQ(n), dK(n) | dP(n+1)
// so that the compute warps' dS(n) phase runs under dV(n) / S(n+1), their P(n+1) phase under dQ(n) / dK(n), and the dQ
// drain under dK(n).
#include "qa_ptx.cuh"
#include "qa_host.h"

namespace qa {

struct Bf16BwdParams2 {
  const float* lse;      // [BH*S] log2 domain
  const float* delta;    // [BH*S]
  const float* dO_f32;   // [BH*S, D] (causal only): row 0 of a head attends uniformly to all keys (LEDGER B-1), dV[k] += dO[0] / S
  int S, causal, BH;
  int S_valid;           // rows [S_valid, S) of every head are zero padding (ragged sequence): padded keys get P = 0
  float sm_scale, qk_scale;
  long long* dbg;        // development library only: [item = head * key tiles + key tile][64] globaltimer stamps (tools/timeline_bf16_bwd.py)
};

#ifdef QA_DEV_TIMELINE
#define QA_TL2(slot)                                                                                      \
  do {                                                                                                    \
    if (p.dbg != nullptr) {                                                                               \
      long long t_;                                                                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                              \
      p.dbg[((size_t)it.bh * nkt + it.j) * 64 + (slot)] = t_;                            \
    }                                                                                                     \
  } while (0)
#else
#define QA_TL2(slot) do { } while (0)
#endif

struct Bf16Bwd2Smem {
  static constexpr int kTile = 128 * 128 * 2;
  static constexpr int off_k = 0;
  static constexpr int off_v = kTile;
  static constexpr int off_q = 2 * kTile;          // 2 stages
  static constexpr int off_do = 4 * kTile;         // 1 stage
  static constexpr int off_ds = 5 * kTile;
  static constexpr int off_st = 6 * kTile;         // output staging: two [128 rows][32 d] fp32 atoms (128 B swizzle)
  static constexpr int off_ld = 7 * kTile;         // [stage][lse | delta][128] fp32
  static constexpr int off_bar = off_ld + 2048;
  static constexpr int off_clc = off_bar + 192;    // two 16-byte cluster-launch-control responses
  static constexpr int used = off_clc + 32;
  static constexpr int total = 232448;             // everything an SM has; the align-up pad must fit in total - used
};

enum Bf16Bwd2Bar { KV_FULL = 0, Q_FULL0, Q_FULL1, DO_FULL, Q_FREE0, Q_FREE1, DO_FREE, S_FULL, P_READY, DP_FULL, DS_READY, DS_FREE,
                   DQ_FULL, DQ_FREE, ACC_FULL, ACC_FREE, CLC_FULL0, CLC_FULL1, CLC_FREE0, CLC_FREE1, kNumBars };

__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// Work item = one CTA of the launch grid (order: bf16_bwd2_item).  A running CTA does not exit after its item: it asks
// the hardware to cancel a CTA that has not started yet (cluster launch control) and processes that CTA's item next, which
// keeps the hardware's dynamic load balance and lets the write-back of one item overlap the loads of the next.
struct Bf16Bwd2Item { int j, bh, i0, nt; };
// Launch order (blockIdx.x = w): qa_group_order with G = 16 for causal heads (key tile 0 is the heaviest), G = 1 otherwise.
__device__ __forceinline__ Bf16Bwd2Item bf16_bwd2_item(int w, int BH, int nkt, int nq, int causal) {
  Bf16Bwd2Item it;
  qa_group_order(w, BH, nkt, causal ? 16 : 1, it.j, it.bh);
  it.i0 = causal ? it.j : 0;                           // query tiles before the diagonal see none of these keys
  it.nt = nq - it.i0;
  return it;
}
__device__ __forceinline__ void clc_try_cancel(uint32_t resp_saddr, uint64_t* bar) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
               ::"r"(resp_saddr), "r"(smem_u32(bar)) : "memory");
}
// true: a pending CTA was cancelled, x is its blockIdx.x
__device__ __forceinline__ bool clc_read(uint32_t resp_saddr, int& x) {
  uint32_t ok, cx = 0, cy = 0;
  asm volatile(
      "{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\t"
      "ld.shared.b128 r, [%3];\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
      "selp.u32 %2, 1, 0, p1;\n\t"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, _, _}, r;\n\t}"
      : "+r"(cx), "+r"(cy), "=r"(ok) : "r"(resp_saddr) : "memory");
  x = (int)cx;
  return ok != 0;
}

__global__ void __launch_bounds__(512, 1)
bf16_bwd_ws_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                   const __grid_constant__ CUtensorMap tm_dq, const __grid_constant__ CUtensorMap tm_dk,
                   const __grid_constant__ CUtensorMap tm_dv, Bf16BwdParams2 p) {
  using L = Bf16Bwd2Smem;
  constexpr int D = 128;
  constexpr int kAtom = 128 * 128;                     // 128 rows x 128 B: one swizzle-atom column (64 x 16 bit or 32 x fp32)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if ((int)(smem - smem_raw) > L::total - L::used) __trap();
  float* ld_s = reinterpret_cast<float*>(smem + L::off_ld);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::off_bar);
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(smem + L::off_bar + kNumBars * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nq = p.S / 128;
  const int nkt = (p.S_valid + 127) / 128;

  if (tid == 0) {
    for (int b = 0; b < kNumBars; ++b)
      mbar_init(&bars[b], (b == P_READY || b == DS_READY) ? 8 : ((b == DQ_FREE || b == ACC_FREE) ? 4 : ((b == CLC_FREE0 || b == CLC_FREE1) ? 13 : 1)));
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc<512>(tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_base_s;
  const uint32_t smem_base = smem_u32(smem);
  constexpr uint32_t tA = 0, tB = 128, tDV = 256, tDK = 384;
  const uint32_t clc_a = smem_base + L::off_clc;
  // Barrier phases run on two counters kept identically by every role: T = tiles so far (all items), I = items so far.
  // After item I every consumer warp reads the response of query I (slot I & 1) and releases the slot.
  auto next_item = [&](uint32_t I, Bf16Bwd2Item& it, bool whole_warp) -> bool {   // whole_warp = false: a single elected thread
    const int s = I & 1;
    int x;
    mbar_wait(&bars[CLC_FULL0 + s], (I >> 1) & 1);
    const bool ok = clc_read(clc_a + s * 16, x);
    fence_proxy_async_smem();                          // the slot is written next through the async proxy
    if (whole_warp) __syncwarp();
    if (!whole_warp || lane == 0) mbar_arrive(&bars[CLC_FREE0 + s]);
    if (ok) it = bf16_bwd2_item(x, p.BH, nkt, nq, p.causal);
    return ok;
  };

  if (warp < 4) {
    // =========================== drain: dQ of every tile, dV / dK of every item (lane = row, column = d) ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int qrow = warp * 32 + lane;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    const float sm = p.sm_scale;
    // each warp stages its own 32 rows ([32 rows][32 d] fp32 = 4 KB of a swizzled atom) and issues its own TMA operation
    auto stage_out = [&](const uint32_t (&r)[128], float sc, const CUtensorMap* tm, int row0, bool reduce, const float* add, float addsc) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t atom = smem_base + L::off_st + (c & 1) * kAtom;
        if (lane == 0) tma_store_wait_read1();         // this warp's operation of two rounds ago has read the atom rows
        __syncwarp();
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (add != nullptr) a4 = __ldg(reinterpret_cast<const float4*>(add + c * 32 + x));
          sts128f(atom + swz128(qrow, x * 4), fmaf(a4.x, addsc, __uint_as_float(r[c * 32 + x]) * sc),
                  fmaf(a4.y, addsc, __uint_as_float(r[c * 32 + x + 1]) * sc), fmaf(a4.z, addsc, __uint_as_float(r[c * 32 + x + 2]) * sc),
                  fmaf(a4.w, addsc, __uint_as_float(r[c * 32 + x + 3]) * sc));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const void* src = smem + L::off_st + (c & 1) * kAtom + warp * 4096;
          if (reduce) tma_reduce_add_2d(tm, src, c * 32, row0); else tma_store_2d(tm, src, c * 32, row0);
          tma_store_commit();
        }
      }
    };
    uint32_t T = 0;
    Bf16Bwd2Item it = bf16_bwd2_item(blockIdx.x, p.BH, nkt, nq, p.causal);
    for (uint32_t I = 0;; ++I) {
      const int head_row0 = it.bh * p.S;
      for (int n = 0; n < it.nt; ++n, ++T) {
        uint32_t r[128];
        mbar_wait(&bars[DQ_FULL], T & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(lane_addr + tB + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
        tmem_ld_wait();
        tc_fence_before();                             // every dQ column is in registers: TMEM B is free for the next dP
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[DQ_FREE]);
        stage_out(r, sm, &tm_dq, head_row0 + (it.i0 + n) * 128 + warp * 32, true, nullptr, 0.f);
      }
      // ---- dV_j, dK_j of the finished item (lane = key)
      mbar_wait(&bars[ACC_FULL], I & 1);
      tc_fence_after();
      if (tid == 0) QA_TL2(11);
      const int k0 = head_row0 + it.j * 128 + warp * 32;
      {
        uint32_t r[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(lane_addr + tDV + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
        tmem_ld_wait();
        // causal: query row 0 of the head has weight 1 / S on every key (LEDGER B-1; its P is masked to 0 in the tiles above)
        stage_out(r, 1.0f, &tm_dv, k0, false, p.causal ? p.dO_f32 + (size_t)head_row0 * D : nullptr, 1.0f / (float)p.S_valid);
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(lane_addr + tDK + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&r[c * 32]));
        tmem_ld_wait();
        tc_fence_before();                             // both accumulators are in registers: the next item may overwrite them
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[ACC_FREE]);
        stage_out(r, sm, &tm_dk, k0, false, nullptr, 0.f);
      }
      if (tid == 0) QA_TL2(12);
      if (!next_item(I, it, true)) break;
    }
    if (lane == 0) tma_store_wait_read();
  } else if (warp < 12) {
    // =========================== compute: P^T (-> TMEM) and dS^T (-> shared memory) ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
    const int cw = warp - 4;
    const int quad = cw & 3, half = cw >> 2;
    const int row = quad * 32 + lane;                  // key inside the tile = TMEM lane
    const uint32_t lane_addr = tbase + ((uint32_t)(quad * 32) << 16);
    const uint32_t ds_base = smem_base + L::off_ds + half * kAtom;
    const float qk = p.qk_scale;
    uint32_t T = 0;
    Bf16Bwd2Item it = bf16_bwd2_item(blockIdx.x, p.BH, nkt, nq, p.causal);
    for (uint32_t I = 0;; ++I) {
      const bool tailk = (it.j + 1) * 128 > p.S_valid;   // this key tile is the ragged last one
      const bool keypad = it.j * 128 + row >= p.S_valid; // this thread's key is padding: P = 0
      for (int n = 0; n < it.nt; ++n, ++T) {
        const int st = T & 1;
        const uint32_t ph = T & 1;
        const bool cdiag = p.causal && (n == 0);       // i == j needs the causal mask
        const bool masked = cdiag || tailk;            // the ragged last key tile needs the key mask
        const uint32_t lse_a = smem_base + L::off_ld + (st * 256 + half * 64) * 4;
        const uint32_t dl_a = lse_a + 512;
        __half2 pk[32];                                // P as packed fp16 (kept for the dS phase)
        mbar_wait(&bars[Q_FULL0 + st], (T >> 1) & 1);  // lse / delta of this tile have landed
        mbar_wait(&bars[S_FULL], ph);
        tc_fence_after();
        {
          uint32_t rs[64];
          tmem_ld32(lane_addr + tA + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&rs[0]));
          tmem_ld32(lane_addr + tA + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&rs[32]));
          tmem_ld_wait();
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {             // 16 queries per TMEM store
            uint32_t w[8];
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              const int x = ch * 16 + c;
              const float4 l4 = lds128f(lse_a + x * 4);
              float e0 = ex2_approx(fmaf(__uint_as_float(rs[x + 0]), qk, -l4.x));     // attention_bf16.py:391-392
              float e1 = ex2_approx(fmaf(__uint_as_float(rs[x + 1]), qk, -l4.y));
              float e2 = ex2_approx(fmaf(__uint_as_float(rs[x + 2]), qk, -l4.z));
              float e3 = ex2_approx(fmaf(__uint_as_float(rs[x + 3]), qk, -l4.w));
              if (masked) {                            // strict causal: key < query keeps its weight (row 0 of the head: fixup kernel)
                const int q0 = cdiag ? half * 64 + x : 0x7fffff00;
                if (row >= q0 + 0 || keypad) e0 = 0.f;
                if (row >= q0 + 1 || keypad) e1 = 0.f;
                if (row >= q0 + 2 || keypad) e2 = 0.f;
                if (row >= q0 + 3 || keypad) e3 = 0.f;
              }
              pk[x / 2 + 0] = __floats2half2_rn(e0, e1);
              pk[x / 2 + 1] = __floats2half2_rn(e2, e3);
              __nv_bfloat162 b0 = __floats2bfloat162_rn(e0, e1), b1 = __floats2bfloat162_rn(e2, e3);
              w[c / 2 + 0] = *reinterpret_cast<uint32_t*>(&b0);
              w[c / 2 + 1] = *reinterpret_cast<uint32_t*>(&b1);
            }
            tmem_st8(lane_addr + tA + half * 64 + ch * 8, w);   // bf16 pairs over the S^T columns this warp has already read
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[P_READY]);

        mbar_wait(&bars[DP_FULL], ph);
        tc_fence_after();
        {
          uint32_t rp[64];
          tmem_ld32(lane_addr + tB + half * 64, *reinterpret_cast<uint32_t(*)[32]>(&rp[0]));
          tmem_ld32(lane_addr + tB + half * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&rp[32]));
          if (T > 0) mbar_wait(&bars[DS_FREE], (T - 1) & 1);     // dQ / dK of the previous tile have read the dS^T buffer
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g) {                // 8 queries = one 16 B store
            uint32_t wd[4];
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              const int x = g * 8 + e * 2;
              const float4 d4 = lds128f(dl_a + x * 4);
              // dS = P * (dP - delta): both factors rounded to fp16, product rounded to fp16 (3 x 2^-12 relative)
              const __half2 t01 = __floats2half2_rn(__uint_as_float(rp[x + 0]) - d4.x, __uint_as_float(rp[x + 1]) - d4.y);
              const __half2 t23 = __floats2half2_rn(__uint_as_float(rp[x + 2]) - d4.z, __uint_as_float(rp[x + 3]) - d4.w);
              __half2 h0 = __hmul2(pk[x / 2 + 0], t01), h1 = __hmul2(pk[x / 2 + 1], t23);
              wd[e + 0] = *reinterpret_cast<uint32_t*>(&h0);
              wd[e + 1] = *reinterpret_cast<uint32_t*>(&h1);
            }
            sts128(ds_base + swz128(row, g * 16), wd[0], wd[1], wd[2], wd[3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[DS_READY]);
      }
      if (!next_item(I, it, true)) break;
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 12) {
      // =========================== tcgen05.mma issue ===========================
      if (elect_one()) {
        constexpr uint32_t id_s = umma_idesc(1, 0, 0, 0, 0, 128, 128);      // S^T : f16, K (K-major) x Q (K-major)
        constexpr uint32_t id_dp = umma_idesc(1, 1, 1, 0, 0, 128, 128);     // dP^T: bf16, V x dO
        constexpr uint32_t id_dv = umma_idesc(1, 1, 1, 0, 1, 128, D);       // dV  : bf16, P^T (TMEM) x dO (MN-major)
        constexpr uint32_t id_dk = umma_idesc(1, 0, 0, 0, 1, 128, D);       // dK  : f16, dS^T (K-major) x Q (MN-major)
        constexpr uint32_t id_dq = umma_idesc(1, 0, 0, 1, 1, 128, D);       // dQ  : f16, dS^T (MN-major: M = query) x K (MN-major)
        const uint32_t a_k = smem_u32(smem + L::off_k), a_v = smem_u32(smem + L::off_v), a_ds = smem_u32(smem + L::off_ds);
        const uint32_t a_do = smem_u32(smem + L::off_do);
        auto issue_s = [&](int st) {
          const uint32_t a_q = smem_u32(smem + L::off_q + st * L::kTile);
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t o = (k >> 2) * kAtom + (k & 3) * 32;
            umma_f16_ss(tbase + tA, umma_smem_desc(a_k + o, 16, 1024, kSwz128), umma_smem_desc(a_q + o, 16, 1024, kSwz128), id_s, k > 0);
          }
          umma_commit(&bars[S_FULL]);
        };
        auto issue_dp = [&]() {
#pragma unroll
          for (int k = 0; k < D / 16; ++k) {
            const uint32_t o = (k >> 2) * kAtom + (k & 3) * 32;
            umma_f16_ss(tbase + tB, umma_smem_desc(a_v + o, 16, 1024, kSwz128), umma_smem_desc(a_do + o, 16, 1024, kSwz128), id_dp, k > 0);
          }
          umma_commit(&bars[DP_FULL]);
        };
        uint32_t T = 0;
        Bf16Bwd2Item it = bf16_bwd2_item(blockIdx.x, p.BH, nkt, nq, p.causal);
        for (uint32_t I = 0;; ++I) {
          QA_TL2(0);
#ifdef QA_DEV_TIMELINE
          if (p.dbg != nullptr) p.dbg[((size_t)it.bh * nkt + it.j) * 64 + 14] = (long long)blockIdx.x + 1;
#endif
          // first tile of the item: S(0) over the P^T of the previous item's last tile (its dV was issued before, in order),
          // dP(0) once the previous item's last dQ has left TMEM B
          mbar_wait(&bars[KV_FULL], I & 1);
          mbar_wait(&bars[Q_FULL0 + (T & 1)], (T >> 1) & 1);
          tc_fence_after();
          QA_TL2(2);
          issue_s(T & 1);
          mbar_wait(&bars[DO_FULL], T & 1);
          if (T > 0) mbar_wait(&bars[DQ_FREE], (T - 1) & 1);
          tc_fence_after();
          issue_dp();
          for (int n = 0; n < it.nt; ++n, ++T) {
            const int st = T & 1;
            const uint32_t a_q = smem_u32(smem + L::off_q + st * L::kTile);
            mbar_wait(&bars[P_READY], T & 1);
            if (n == 0 && I > 0) mbar_wait(&bars[ACC_FREE], (I - 1) & 1);   // the drain warps hold dV / dK of the previous item
            tc_fence_after();
            if (n == 0) QA_TL2(3);
#pragma unroll
            for (int k = 0; k < 8; ++k)                           // contraction over the 128 queries, 16 per instruction
              umma_f16_ts(tbase + tDV, tbase + tA + (k >> 2) * 64 + (k & 3) * 8, umma_smem_desc(a_do + k * 2048, kAtom, 1024, kSwz128),
                          id_dv, (n > 0) || (k > 0));
            umma_commit(&bars[DO_FREE]);
            if (n + 1 < it.nt) {
              mbar_wait(&bars[Q_FULL0 + (st ^ 1)], ((T + 1) >> 1) & 1);
              tc_fence_after();
              issue_s(st ^ 1);                                    // overwrites P^T(n): tcgen05.mma of one thread execute in order
            }
            mbar_wait(&bars[DS_READY], T & 1);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 8; ++k)                           // contraction over the 128 keys
              umma_f16_ss(tbase + tB, umma_smem_desc(a_ds + k * 2048, kAtom, 1024, kSwz128),
                          umma_smem_desc(a_k + k * 2048, kAtom, 1024, kSwz128), id_dq, k > 0);
            umma_commit(&bars[DQ_FULL]);
#pragma unroll
            for (int k = 0; k < 8; ++k)                           // contraction over the 128 queries
              umma_f16_ss(tbase + tDK, umma_smem_desc(a_ds + (k >> 2) * kAtom + (k & 3) * 32, 16, 1024, kSwz128),
                          umma_smem_desc(a_q + k * 2048, kAtom, 1024, kSwz128), id_dk, (n > 0) || (k > 0));
            umma_commit(&bars[Q_FREE0 + st]);
            umma_commit(&bars[DS_FREE]);
            if (n + 1 < it.nt) {
              mbar_wait(&bars[DO_FULL], (T + 1) & 1);
              mbar_wait(&bars[DQ_FREE], T & 1);
              tc_fence_after();
              issue_dp();
            }
          }
          umma_commit(&bars[ACC_FULL]);                           // every MMA of the item is complete: K / V may be replaced
          if (!next_item(I, it, false)) break;
        }
      }
      __syncwarp();
    } else if (warp == 13) {
      // =========================== TMA loads ===========================
      if (elect_one()) {
        uint32_t T = 0;
        Bf16Bwd2Item it = bf16_bwd2_item(blockIdx.x, p.BH, nkt, nq, p.causal);
        for (uint32_t I = 0;; ++I) {
          const int head_row0 = it.bh * p.S;
          {                                                       // query I: claim the item after this one
            const int s = I & 1;
            if (I >= 2) mbar_wait(&bars[CLC_FREE0 + s], ((I >> 1) - 1) & 1);
            mbar_expect_tx(&bars[CLC_FULL0 + s], 16);
            clc_try_cancel(clc_a + s * 16, &bars[CLC_FULL0 + s]);
          }
          for (int n = 0; n < it.nt; ++n, ++T) {
            const int st = T & 1;
            const int r0 = head_row0 + (it.i0 + n) * 128;
            if (T >= 2) mbar_wait(&bars[Q_FREE0 + st], ((T >> 1) - 1) & 1);
            mbar_expect_tx(&bars[Q_FULL0 + st], L::kTile + 1024);
#pragma unroll
            for (int a = 0; a < 2; ++a) tma_load_2d(smem + L::off_q + st * L::kTile + a * kAtom, &tm_q, &bars[Q_FULL0 + st], a * 64, r0);
            bulk_load_1d(ld_s + st * 256, p.lse + r0, 512, &bars[Q_FULL0 + st]);
            bulk_load_1d(ld_s + st * 256 + 128, p.delta + r0, 512, &bars[Q_FULL0 + st]);
            if (T >= 1) mbar_wait(&bars[DO_FREE], (T - 1) & 1);  // dV of the previous tile, the last reader of the dO tile
            mbar_expect_tx(&bars[DO_FULL], L::kTile);
#pragma unroll
            for (int a = 0; a < 2; ++a) tma_load_2d(smem + L::off_do + a * kAtom, &tm_do, &bars[DO_FULL], a * 64, r0);
            if (n == 0) {                                         // K / V after the first Q / dO tile: they wait for the previous item
              if (I > 0) mbar_wait(&bars[ACC_FULL], (I - 1) & 1);
              mbar_expect_tx(&bars[KV_FULL], 2 * L::kTile);
#pragma unroll
              for (int a = 0; a < 2; ++a) {
                tma_load_2d(smem + L::off_k + a * kAtom, &tm_k, &bars[KV_FULL], a * 64, head_row0 + it.j * 128);
                tma_load_2d(smem + L::off_v + a * kAtom, &tm_v, &bars[KV_FULL], a * 64, head_row0 + it.j * 128);
              }
            }
          }
          int x;
          mbar_wait(&bars[CLC_FULL0 + (I & 1)], (I >> 1) & 1);
          if (!clc_read(clc_a + (I & 1) * 16, x)) break;
          it = bf16_bwd2_item(x, p.BH, nkt, nq, p.causal);
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) tmem_dealloc<512>(tbase);
}

#ifdef QA_DEV_TIMELINE
static void* g_bf16_bwd_dbg = nullptr;
// Development library only (include/qattn_dev.h, tools/timeline_bf16_bwd.py): per-CTA globaltimer stamps of subsequent
// qa_bf16_bwd launches ([CTAs][64] int64); NULL = off.
extern "C" int qa_debug_set_bf16_bwd_timeline(void* buf) {
  g_bf16_bwd_dbg = buf;
  return 0;
}
#endif

int launch_bf16_bwd_ws(const void* q, const void* k, const void* v, const void* do_bf16, const float* dO_f32, const float* lse,
                       const float* delta, float* dq, float* dk, float* dv, int BH, int S, int S_valid, int causal, cudaStream_t st) {
  using L = Bf16Bwd2Smem;
  constexpr int D = 128;
  CUtensorMap tq, tk, tv, tdo, tdq, tdk, tdv;
  uint64_t dims[2] = {(uint64_t)D, (uint64_t)BH * S};
  uint64_t str[1] = {(uint64_t)D * 2};
  uint32_t box[2] = {64, 128};
  uint64_t str32[1] = {(uint64_t)D * 4};
  uint32_t boxq[2] = {32, 32};                        // dQ / dK / dV: one drain warp's 32 rows x 32 columns
  int rc;
  if ((rc = qa_make_tmap(&tq, q, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tk, k, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tv, v, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tdo, do_bf16, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, dims, str, box, 3))) return rc;
  if ((rc = qa_make_tmap(&tdq, dq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, str32, boxq, 3))) return rc;
  if ((rc = qa_make_tmap(&tdk, dk, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, str32, boxq, 3))) return rc;
  if ((rc = qa_make_tmap(&tdv, dv, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dims, str32, boxq, 3))) return rc;
  Bf16BwdParams2 p;
  p.lse = lse; p.delta = delta; p.dO_f32 = dO_f32; p.S = S; p.S_valid = S_valid; p.causal = causal; p.BH = BH;
  p.sm_scale = (float)(1.0 / sqrt((double)D));
  p.qk_scale = (float)((1.0 / sqrt((double)D)) * 1.44269504);
#ifdef QA_DEV_TIMELINE
  p.dbg = (long long*)g_bf16_bwd_dbg;
#else
  p.dbg = nullptr;
#endif
  cudaError_t e = cudaFuncSetAttribute(bf16_bwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total);
  if (e != cudaSuccess) return qa_fail(QA_ERR_CUDA, cudaGetErrorString(e));
  dim3 grid(((S_valid + 127) / 128) * BH);             // one CTA per item; running CTAs take over the items of pending ones
  bf16_bwd_ws_kernel<<<grid, 512, L::total, st>>>(tq, tk, tv, tdo, tdq, tdk, tdv, p);
  return qa_check_launch("qa_bf16_bwd(ws)");
}

}  // namespace qa
