/* qattn.h -- C ABI of libqattn.so: B200 (sm_100a) kernels for the three fused attention paths of
 * selau642/QuantizedAttention.  This is the drop-in boundary: the reference's Python kernels are replaced by
 * calls into these entry points (ctypes binding: quantizedattention_b200/_lib.py).
 *
 * Conventions
 *   - Plain pointers and sizes only.  The CALLER owns every buffer (inputs, outputs, workspaces are device
 *     pointers, e.g. torch `tensor.data_ptr()`); the library never allocates or frees device memory and never
 *     synchronises.  All work is enqueued on `stream` (a cudaStream_t passed as void*) of the current device.
 *   - Tensors are row-major [B*H*S, D] ("[N, D]") views of the reference's contiguous [B, H, S, D] layout.
 *   - Return value: 0 on success, negative QA_ERR_* otherwise; qa_last_error() gives the message (thread-local).
 *   - D in {64, 128}; sequence lengths multiples of 128; 16-byte aligned pointers (TMA).
 *   - Logits live in the log2 domain: qk_scale = log2(e)/sqrt(D); "lse" is log2-sum-exp2 (attention_int8.py:252).
 */
#ifndef QATTN_H
#define QATTN_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define QA_ERR_SHAPE (-1)
#define QA_ERR_ALIGN (-2)
#define QA_ERR_WORKSPACE (-3)
#define QA_ERR_CUDA (-4)
#define QA_ERR_ARCH (-5)
#define QA_ERR_DRIVER (-6)

int qa_version(void);
const char* qa_last_error(void);

/* Bytes of the caller-owned workspace an op needs (the library never allocates device memory). */
#define QA_WS_K_MEAN 0             /* qa_k_mean / qa_k_token_sum `workspace` */
#define QA_WS_INT8_BWD_DQ 1        /* qa_int8_bwd `dq_ws_f32` (zero-initialised by the caller) */
#define QA_WS_INT8_BWD_ROWSUM 2    /* qa_int8_bwd `rowsum_ws_f32` (zero-initialised by the caller) */
#define QA_WS_JVP_BF16_OPERANDS 3  /* the six bf16 operand copies qa_jvp_fwd reads (made with qa_cast_f32) */
size_t qa_workspace_bytes(int op, int B, int H, int S, int D);

/* ---- int8 path pre-passes (replace the per-tile-pair recomputation inside attention_int8.py:178-195, 241-247) ---- */

/* K-smoothing mean, attention_int8.py:24-25 under LEDGER I-1: mean over tokens per (b,h), fp32 accumulate, fp16 out
 * [B*H, D].  workspace >= qa_k_mean_workspace_bytes(). */
size_t qa_k_mean_workspace_bytes(int B, int H, int S, int D);
int qa_k_mean(const void* k_fp16, void* mean_fp16, void* workspace, size_t ws_bytes, int B, int H, int S, int D,
              void* stream);

/* fp32 token sums [B*H, D] of a sequence shard: the ring-KV path all-reduces these across ranks, divides by the global
 * sequence length and rounds to fp16, so that every rank smooths K with the same mean. */
int qa_k_token_sum(const void* k_fp16, void* sum_f32, void* workspace, size_t ws_bytes, int B, int H, int S, int D,
                   void* stream);

/* Per-block int8 quantisation, attention_int8.py:178-186 (Q), :188-195 (K), :241-247 (V), :369-374 (dO):
 * scale = fp16(amax|block| / 127); value = trunc(fp16(x / scale)); block = blk rows x D.  Bit-exact with the reference
 * arithmetic.  mean_fp16 != NULL subtracts the per-head mean first (fp16, one rounding): K-smoothing fused in.
 * rounding: 2 = fp8: e4m3 codes with scale = fp16(amax / 448) (qa_fp8_fwd).
 * rounding: 0 = truncate toward zero, the reference's `.to(torch.int8)` (attention_int8.py:183); 1 = round half to even,
 * the opt-in accuracy mode (removes the truncation bias; not bit-comparable with the reference by construction).
 * qa_int8_fwd / qa_int8_bwd take it as bit 0 of their `flags` (QA_FLAG_NEAREST); bit 1 (QA_FLAG_CAUSAL) selects the
 * strict causal mask of the reference's baseline (key < query, attention_int8.py:465-473; row 0 of a head = uniform
 * average over all keys), a mode the reference's int8 kernel does not have (Sq == Sk, Bq = Bkv = 128). */
#define QA_FLAG_NEAREST 1
#define QA_FLAG_CAUSAL 2
#define QA_FLAG_BWD_8WARP 4   /* qa_int8_bwd only: the 8-warp (not warp-specialised) kernel; default = warp-specialised */
int qa_quant_block(const void* x_fp16, const void* mean_fp16, void* out_i8, void* scales_fp16, long long n_rows, int D,
                   int blk, int rows_per_head, int rounding, void* stream);

/* ---- int8 forward: helion_atten_int8_hl_dot_fwd, attention_int8.py:101-262 (tile loop :170-257) ----
 * Normal mode: O fp16 [BH*Sq, D], lse16 fp16, lse32 fp32 (optional).  Ring mode (o_acc != NULL): unnormalised fp32
 * accumulator plus running (m, l) per row for the K/V shard, merged by the caller (sequence-sharded ring KV). */
int qa_int8_fwd(const void* q_i8, const void* k_i8, const void* v_i8, const void* sq_fp16, const void* sk_fp16,
                const void* sv_fp16, void* O_fp16, void* lse_fp16, void* lse_fp32, void* o_acc_fp32, void* m_out_fp32,
                void* l_out_fp32, int BH, int Sq, int Sk, int D, int Bq, int Bkv, int nsplit, int flags, void* stream);

/* Same kernel, continuing from a running online-softmax state (o_acc_in, m_in, l_in) produced by earlier K/V shards of
 * the ring; all three NULL = fresh state.  In/out state buffers may alias. */
int qa_int8_fwd_state(const void* q_i8, const void* k_i8, const void* v_i8, const void* sq_fp16, const void* sk_fp16,
                      const void* sv_fp16, void* O_fp16, void* lse_fp16, void* lse_fp32, void* o_acc_fp32, void* m_out_fp32,
                      void* l_out_fp32, const void* o_acc_in_fp32, const void* m_in_fp32, const void* l_in_fp32, int BH,
                      int Sq, int Sk, int D, int Bq, int Bkv, int nsplit, int flags, void* stream);

/* Ragged sequences (the reference's hl.tile clamps the last tile, attention_int8.py:170,176): the buffers are zero-padded
 * per head to Sk (a multiple of 128 and of Bkv), keys [Sk_valid, Sk) have weight exactly 0 and k-tiles without a valid key
 * are skipped.  qa_int8_fwd_state == qa_int8_fwd_ragged with Sk_valid = Sk.  Non-causal only. */
int qa_int8_fwd_ragged(const void* q_i8, const void* k_i8, const void* v_i8, const void* sq_fp16, const void* sk_fp16,
                       const void* sv_fp16, void* O_fp16, void* lse_fp16, void* lse_fp32, void* o_acc_fp32, void* m_out_fp32,
                       void* l_out_fp32, const void* o_acc_in_fp32, const void* m_in_fp32, const void* l_in_fp32, int BH,
                       int Sq, int Sk, int Sk_valid, int D, int Bq, int Bkv, int nsplit, int flags, void* stream);

/* ---- fp8 (e4m3) forward, SURVEY.md 8f.4 (the reference names it, README.md:48-54, but ships no code): the int8
 * pipeline with e4m3 operands (qa_quant_block rounding = 2: scale = amax / 448), tcgen05 kind::f8f6f4, fp32 TMEM
 * accumulators; Bq = Bkv = 128.  O fp16 [BH*Sq, D], lse fp16 / fp32 [BH*Sq] (lse32 optional). */
int qa_fp8_fwd(const void* q_e4m3, const void* k_e4m3, const void* v_e4m3, const void* sq_fp16, const void* sk_fp16,
               const void* sv_fp16, void* O_fp16, void* lse_fp16, void* lse_fp32, int BH, int Sq, int Sk, int D, void* stream);

/* ---- NVFP4 (microscaling) forward, SURVEY.md 8f.4 (README.md:48-54 of the reference names FP4 microscaling as the
 * SageAttention3 feature; it ships no code for it): tcgen05 kind::mxf4nvf4.block_scale.block16 for both contractions.
 * Two-level scales: sg = amax_head / 2688 (fp32 [BH]), sf = e4m3(amax_blk16 / 6 / sg), code = e2m1_rn(x / (sf * sg));
 * D = 128, S % 128 == 0.  Q, K: blocks of 16 along D, codes [BH*S, D/2] bytes (element 2i in the low nibble); V: blocks of
 * 16 KEYS, codes transposed [BH, D, S/2].  Scale factors in the tcgen05.cp atom layout: per 128-row tile and 64-element
 * K step 512 bytes, byte 16*(r%32) + 4*(r/32) + s = row r, block s.  amax_ws: 2*BH 32-bit words of scratch (the head amax
 * is formed inside the one-pass kernel: atomicMax + arrival counter per head).  mean_fp16 (or NULL):
 * per-head K token mean subtracted first (one fp16 rounding), as in the int8 path. */
int qa_fp4_quant_rows(const void* x_fp16, const void* mean_fp16, void* amax_ws, void* codes, void* sf, void* sg_f32, int BH, int S,
                      int D, void* stream);
/* Ragged sequences (zero-padded per head to S, a multiple of 128): rows [S_valid, S) of every head stay zero after the smoothing
 * and do not enter the head amax; qa_fp4_fwd_ragged gives the keys [Sk_valid, Sk) weight exactly 0 (variant 0). */
int qa_fp4_quant_rows_ragged(const void* x_fp16, const void* mean_fp16, void* amax_ws, void* codes, void* sf, void* sg_f32, int BH,
                             int S, int S_valid, int D, void* stream);
int qa_fp4_quant_vt(const void* v_fp16, void* amax_ws, void* codes_t, void* sf, void* sg_f32, int BH, int S, int D, void* stream);
/* O fp16 [BH*Sq, D], lse fp32 [BH*Sq] (log2 domain); P is microscaled per row and 16 keys inside the kernel
 * (sfp = e4m3(amax * 448), code = e2m1_rn(P * 2688 / sfp)); the fp32 accumulator spans all k-tiles.
 * variant 0 (default): one CTA per SM, 128-key tiles, running-maximum warps a tile ahead of two alternating exp warps per
 * row group; 1: two CTAs per SM, 64-key steps (same numerics up to the step size of the online softmax).
 * flags: QA_FLAG_CAUSAL (variant 0, Sq == Sk) = the strict mask of the reference's baseline, as on the int8 path: key < query,
 * row 0 of a head = uniform average over all keys of the de-quantised V, lse = -128 + log2(S). */
int qa_fp4_fwd(const void* q4, const void* sfq, const void* sgq_f32, const void* k4, const void* sfk, const void* sgk_f32,
               const void* vt4, const void* sfv, const void* sgv_f32, void* O_fp16, void* lse_f32, int BH, int Sq, int Sk, int D,
               int variant, int flags, void* stream);
int qa_fp4_fwd_ragged(const void* q4, const void* sfq, const void* sgq_f32, const void* k4, const void* sfk, const void* sgk_f32,
                      const void* vt4, const void* sfv, const void* sgv_f32, void* O_fp16, void* lse_f32, int BH, int Sq, int Sk,
                      int Sk_valid, int D, int variant, int flags, float sm_scale /* <= 0: 1/sqrt(D) */, void* stream);

/* ---- backward pre/post passes ---- */
/* delta = rowsum(dO * O) fp32 (attention_int8.py:397-398, attention_bf16.py:416).  in_dtype 0: fp16 dO/O;
 * 1: fp32 dO/O and, if dO_bf16 != NULL, a bf16 copy of dO in the same pass. */
int qa_bwd_delta(const void* dO, const void* O, void* delta_f32, void* dO_bf16, long long n_rows, int D, int in_dtype,
                 void* stream);
/* fp32 -> fp16 (out_dtype 0) / bf16 (1) */
int qa_cast_f32(const void* in_f32, void* out, long long n, int out_dtype, void* stream);

/* ---- int8 backward: helion_atten_int8_hl_dot_bwd, attention_int8.py:268-432 under the 8-LEDGER contract ----
 * Bq, Bkv in {32, 64, 128} (causal: 128 / 128).  dq_ws: zero-initialised fp32 [BH*S, D] accumulator; rowsum_ws: zero-initialised fp32 [BH*S]
 * accumulator of rowsum(dS) for the K-smoothing term (NULL when K was not smoothed); dk, dv fp16.
 * qa_int8_bwd_finalize turns the two workspaces into dq = fp16(dq_ws + sm_scale * rowsum * k_mean[b,h]). */
int qa_int8_bwd(const void* q_i8, const void* k_i8, const void* v_i8, const void* do_i8, const void* sq_fp16,
                const void* sk_fp16, const void* sv_fp16, const void* s_do_fp16, const void* lse_f32,
                const void* delta_f32, void* rowsum_ws_f32, void* dq_ws_f32, void* dk_f16, void* dv_f16, int BH, int S,
                int D, int Bq, int Bkv, int flags, void* stream);
/* Same for a ragged sequence padded to S: rows [S_valid, S) of every head are padding (dO rows zero); padded keys get
 * P = 0, k-tiles and query tiles without a valid row are skipped (their dk / dv rows are not written). */
int qa_int8_bwd_ragged(const void* q_i8, const void* k_i8, const void* v_i8, const void* do_i8, const void* sq_fp16,
                       const void* sk_fp16, const void* sv_fp16, const void* s_do_fp16, const void* lse_f32,
                       const void* delta_f32, void* rowsum_ws_f32, void* dq_ws_f32, void* dk_f16, void* dv_f16, int BH, int S,
                       int S_valid, int D, int Bq, int Bkv, int flags, void* stream);
/* SageBwd option: dP = dO V^T from the UNQUANTISED fp16 dO and V (tcgen05 kind::f16, fp32 accumulation)
 * instead of the int8 product the reference uses (attention_int8.py:380-384); the other four contractions stay int8.
 * Bq, Bkv in {32, 64, 128}; non-causal; S a multiple of 128.  flags: QA_FLAG_NEAREST only. */
int qa_int8_bwd_sage(const void* q_i8, const void* k_i8, const void* v_fp16, const void* do_i8, const void* do_fp16,
                     const void* sq_fp16, const void* sk_fp16, const void* s_do_fp16, const void* lse_f32, const void* delta_f32,
                     void* rowsum_ws_f32, void* dq_ws_f32, void* dk_f16, void* dv_f16, int BH, int S, int D, int Bq, int Bkv,
                     int flags, void* stream);
int qa_int8_bwd_finalize(const void* dq_ws_f32, const void* rowsum_ws_f32, const void* k_mean_f16, void* dq_f16, int BH,
                         int S, int D, void* stream);

/* ---- bf16 path: helion_atten_bf16_fwd_training (attention_bf16.py:111-296), helion_flash_atten_2_algo_4_bwd
 * (attention_bf16.py:309-448) ---- */
int qa_bf16_fwd(const void* q_f16, const void* k_f16, const void* v_bf16, void* O_f32, void* lse_f32, int BH, int Sq,
                int Sk, int D, int causal, int nsplit, void* stream);
/* Same, with the lazy-rescale threshold explicit: a new running maximum m' (attention_bf16.py:236-264) is adopted only
 * when it exceeds the current one by more than rescale_tau (log2 units, 0..16); until then P = exp2(u - m) may reach
 * 2^rescale_tau and O is not rescaled (:280).  Mathematically neutral; 0 reproduces the reference's step-by-step
 * maximum, qa_bf16_fwd uses 8. */
int qa_bf16_fwd_ex(const void* q_f16, const void* k_f16, const void* v_bf16, void* O_f32, void* lse_f32, int BH, int Sq,
                   int Sk, int D, int causal, int nsplit, float rescale_tau, void* stream);
/* Ragged sequences (the reference's hl.tile clamps the last tile, attention_bf16.py:170,201): buffers zero-padded per head
 * to Sq / Sk (multiples of 128); keys [Sk_valid, Sk) have weight exactly 0; Sk - 128 < Sk_valid <= Sk. */
int qa_bf16_fwd_ragged(const void* q_f16, const void* k_f16, const void* v_bf16, void* O_f32, void* lse_f32, int BH, int Sq,
                       int Sk, int Sk_valid, int D, int causal, int nsplit, float rescale_tau, void* stream);
int qa_bf16_bwd(const void* q_f16, const void* k_f16, const void* v_bf16, const void* dO_bf16, const void* dO_f32,
                const void* lse_f32, const void* delta_f32, void* dq_f32, void* dk_f32, void* dv_f32, int BH, int S, int D,
                int causal, void* stream);
/* Same with the kernel explicit.  variant 0 (what qa_bf16_bwd runs): D = 128 -> the warp-specialised kernel (logits
 * computed transposed, P fed to the tensor core from TMEM, dQ^T drained with coalesced reductions); D = 64 -> the
 * phase-sequential kernel.  variant 1: the phase-sequential kernel for every D. */
int qa_bf16_bwd_ex(const void* q_f16, const void* k_f16, const void* v_bf16, const void* dO_bf16, const void* dO_f32,
                   const void* lse_f32, const void* delta_f32, void* dq_f32, void* dk_f32, void* dv_f32, int BH, int S, int D,
                   int causal, int variant, void* stream);

/* Same for a ragged sequence padded to S: rows [S_valid, S) are padding.  The caller pads dO (and O) with zeros and lse with a
 * large finite value, so padded query rows get P = 0; padded keys get P = 0 in the kernel. */
int qa_bf16_bwd_ragged(const void* q_f16, const void* k_f16, const void* v_bf16, const void* dO_bf16, const void* dO_f32,
                       const void* lse_f32, const void* delta_f32, void* dq_f32, void* dk_f32, void* dv_f32, int BH, int S,
                       int S_valid, int D, int causal, int variant, void* stream);

/* ---- JVP: helion_attention_jvp_forward_fp32, attention_jvp.py:33-195 (operands pre-cast to bf16); D in {64,128} ---- */
int qa_jvp_fwd(const void* q_bf16, const void* tq_bf16, const void* k_bf16, const void* tk_bf16, const void* v_bf16,
               const void* tv_bf16, void* O_f32, void* tO_f32, void* lse_f32, int BH, int Sq, int Sk, int D, int nsplit,
               void* stream);

/* Ragged sequences (attention_jvp.py:120,137): buffers zero-padded per head to Sq / Sk (multiples of 128), keys
 * [Sk_valid, Sk) have weight exactly 0. */
int qa_jvp_fwd_ragged(const void* q_bf16, const void* tq_bf16, const void* k_bf16, const void* tk_bf16, const void* v_bf16,
                      const void* tv_bf16, void* O_f32, void* tO_f32, void* lse_f32, int BH, int Sq, int Sk, int Sk_valid, int D,
                      int nsplit, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QATTN_H */
