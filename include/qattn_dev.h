/* qattn_dev.h -- development-only entry points of libqattn_dev.so (built by `python -m quantizedattention_b200.build
 * --dev` with -DQA_DEV_TIMELINE): hardware layout probes (tests/test_probe_gpu.py, tools/tmem_bw_probe.py) and the
 * kernel timeline hooks (tools/timeline*.py).  The product library libqattn.so exports none of these and its kernels
 * carry no debug hooks; the development library additionally exports everything in qattn.h. */
#ifndef QATTN_DEV_H
#define QATTN_DEV_H
#include "qattn.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- hardware probes used by tests/test_probe_gpu.py (layout / descriptor conventions) ---- */
int qa_probe_mma(const void* a_img, int a_bytes, const void* b_img, int b_bytes, void* d_out, int a_lbo, int a_sbo,
                 int a_layout, int a_kstep_bytes, int b_lbo, int b_sbo, int b_layout, int b_kstep_bytes, unsigned idesc,
                 int kind, int n_mma, int n_cols, int a_in_tmem, int a_tmem_cols, int a_tmem_kstep_cols, void* stream);
/* block-scaled (microscaling) variant: packed e2m1 operands, scale factors staged shared memory -> TMEM by tcgen05.cp.
 * kind: 0 = mxf4nvf4.block16 (ue4m3), 1 = mxf4.block32 (ue8m0), 2 = mxf8f6f4.block32 */
int qa_probe_mma_bs(const void* a_img, int a_bytes, const void* b_img, int b_bytes, const void* sfa_img, int sfa_bytes,
                    const void* sfb_img, int sfb_bytes, void* d_out, int a_lbo, int a_sbo, int a_layout, int a_kstep_bytes,
                    int b_lbo, int b_sbo, int b_layout, int b_kstep_bytes, unsigned idesc, int kind, int n_mma, int n_cols,
                    int sfa_cols_per_mma, int sfb_cols_per_mma, int a_in_tmem, int a_tmem_cols, int a_tmem_kstep_cols,
                    void* stream);
/* Development aid: per-k-tile SM-clock stamps of CTA (0,0) of the next qa_int8_fwd launches ([64][16] int64); NULL = off */
int qa_debug_set_int8_fwd_timeline(void* buf);
int qa_debug_set_int8_bwd_timeline(void* buf_i64_64x2x16);   /* same for qa_int8_bwd: leader warp and warp 5, per q-tile */
/* per-item globaltimer stamps ([head * key tiles + key tile][64] int64) of the next qa_bf16_bwd launches (D = 128 kernel) */
int qa_debug_set_bf16_bwd_timeline(void* buf);
/* per-CTA globaltimer stamps ([CTAs][16] int64, slot 15 = SM id) of the next qa_bf16_fwd launches (two-tile kernel) */
int qa_debug_set_bf16_fwd_timeline(void* buf);
/* TMEM -> register read bandwidth (tcgen05.ld.32x32b.x32 streamed by every warp): measured ceiling of the drains */
int qa_probe_tmem_bw(void* sink, int blocks, int threads, int iters, void* stream);
/* shape: 0 = 32x32b.x32, 1 = 16x256b.x8, 2 = 16x128b.x16, 3 = 16x64b.x32; depth = loads in flight per warp (1, 2) */
int qa_probe_tmem_bw_ex(void* sink, int blocks, int threads, int iters, int shape, int depth, void* stream);
int qa_probe_tma(const void* gptr, int elem_bytes, int rank, const unsigned long long* dims,
                 const unsigned long long* strides_bytes, const unsigned* box, int swizzle, const int* coords, void* out,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QATTN_DEV_H */
