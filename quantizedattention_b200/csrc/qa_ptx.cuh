// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA, tcgen05 (MMA / TMEM), proxy fences.
// Everything here is hand-written against the PTX ISA; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace qa {

#ifndef QA_MBAR_HINT_NS
#define QA_MBAR_HINT_NS 0x989680u   // suspend-time hint of mbarrier.try_wait (ns)
#endif
#ifndef QA_SPIN_LIMIT
#define QA_SPIN_LIMIT (1u << 14)   // bounded mbarrier waits (each up to ~10 ms): a protocol bug traps instead of hanging
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint expires)
// instead of polling, so waiting warps do not steal issue slots from the working warps of the same SM sub-partition.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(QA_MBAR_HINT_NS)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > QA_SPIN_LIMIT) { __trap(); }     // a protocol bug traps instead of hanging the GPU
  }
}

// ---------------------------------------------------------------- shared-memory stores
// The kernels' tile pointers are derived from a manually aligned base, which hides the address space from the
// compiler (it then emits generic ST.E with 64-bit addresses); these wrappers keep the stores as STS.128.
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t saddr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------- fences
__device__ __forceinline__ void fence_proxy_async_smem() {   // generic-proxy smem writes -> visible to TMA / tcgen05
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
// TMA reduce-add (element type from the tensor map): global[tile] += smem[tile]
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM alloc
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {        // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (sm_100 format): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), base_offset [49,52), layout type [61,64) (0 none, 2 = 128B, 4 = 64B, 6 = 32B swizzle).
__host__ __device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                            uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type & 7) << 61;
  return d;
}
constexpr uint32_t kSwz128 = 2, kSwz64 = 4, kSwz32 = 6, kSwzNone = 0;

// Instruction descriptor (upper 32 bits of idescE).  Formats: kind::f16 a/b 0=f16 1=bf16; kind::tf32 2;
// kind::i8 0=u8 1=s8.  c_format 0=f16 1=f32 2=s32.  major: 0 = K-major, 1 = MN-major.
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t c_fmt, uint32_t a_fmt, uint32_t b_fmt, uint32_t a_major,
                                                  uint32_t b_major, uint32_t M, uint32_t N) {
  return (c_fmt << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_major << 15) | (b_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

// ---------------------------------------------------------------- tcgen05.mma (single thread issues)
__device__ __forceinline__ void umma_i8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// kind::f8f6f4 (a_fmt / b_fmt in the instruction descriptor: 0 = e4m3, 1 = e5m2), fp32 accumulation; 8-bit operands use
// the same shared-memory / TMEM operand layouts as kind::i8 (one byte per element, K = 32 per instruction)
__device__ __forceinline__ void umma_f8_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_f8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// Block-scaled kinds (microscaling): the scale factors of A and B live in TMEM (copied there with tcgen05.cp); idesc is
// the block-scaled instruction descriptor (umma_idesc_bs).  mxf4nvf4.block16: e2m1 x e2m1, ue4m3 scale per 16 elements,
// K = 64; mxf4.block32: ue8m0 scale per 32 elements, K = 64; mxf8f6f4.block32: K = 32.
#define QA_UMMA_BS(NAME, KIND, AOP, ACON, ATYPE)                                                                                  \
  __device__ __forceinline__ void NAME(uint32_t d_tmem, ATYPE a, uint64_t b_desc, uint32_t idesc, uint32_t sfa_tmem,             \
                                       uint32_t sfb_tmem, uint32_t accum) {                                                      \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                                               \
                 "tcgen05.mma.cta_group::1.kind::" KIND " [%0], " AOP ", %2, %3, [%5], [%6], p;\n\t}"                             \
                 ::"r"(d_tmem), ACON(a), "l"(b_desc), "r"(idesc), "r"(accum), "r"(sfa_tmem), "r"(sfb_tmem)                       \
                 : "memory");                                                                                                     \
  }
QA_UMMA_BS(umma_nvf4_ss, "mxf4nvf4.block_scale.block16", "%1", "l", uint64_t)
QA_UMMA_BS(umma_nvf4_ts, "mxf4nvf4.block_scale.block16", "[%1]", "r", uint32_t)
QA_UMMA_BS(umma_mxf4_ss, "mxf4.block_scale.block32", "%1", "l", uint64_t)
QA_UMMA_BS(umma_mxf8_ss, "mxf8f6f4.block_scale", "%1", "l", uint64_t)
#undef QA_UMMA_BS
// Block-scaled instruction descriptor: formats for the mxf4 kinds: e2m1 = 1 (mxf8f6f4: e4m3 0, e5m2 1, e2m1 5);
// sf_fmt 0 = ue4m3, 1 = ue8m0; a_sf_id / b_sf_id select the byte of the 32-bit scale column (0 when all four are used).
__host__ __device__ constexpr uint32_t umma_idesc_bs(uint32_t a_fmt, uint32_t b_fmt, uint32_t a_major, uint32_t b_major, uint32_t M,
                                                     uint32_t N, uint32_t sf_fmt, uint32_t a_sf_id = 0, uint32_t b_sf_id = 0) {
  return (b_sf_id << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_major << 15) | (b_major << 16) | ((N >> 3) << 17) | (sf_fmt << 23) |
         ((M >> 4) << 24) | (a_sf_id << 29);
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- TMEM <-> registers (32x32b: thread t of warp w owns lane 32*(w%4)+t)
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// 64 consecutive columns as two x32 loads in flight (no wait).  tcgen05.wait::ld is the expensive part of a TMEM read
// (tools/tmem_bw_probe.py: one load per wait ~100 B/clk/SM, two loads per wait 350-450 B/clk/SM), so kernels batch
// as many loads as their registers allow in front of a single wait.
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
  tmem_ld32(taddr, *reinterpret_cast<uint32_t (*)[32]>(&r[0]));
  tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t (*)[32]>(&r[32]));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
      "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
      "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

// ---------------------------------------------------------------- misc
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
// low bytes of four 32-bit words -> one word (w0 lowest): 3 PRMTs
__device__ __forceinline__ uint32_t pack_low_bytes(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
  return prmt(prmt(w0, w1, 0x0040u), prmt(w2, w3, 0x0040u), 0x5410u);
}
// ---------------------------------------------------------------- "magic" accumulators (int8 kernels)
// A kind::f16 MMA over two constant shared-memory tiles (every fp16 element of A = 1024, of B = 768, K = 16) initialises
// an fp32 accumulator to kMagic = 16 * 1024 * 768 = 1.5 * 2^23 = 12582912.0f (bits 0x4B400000), a value of the binade
// [2^23, 2^24) where one unit in the last place is 1.  The kind::i8 MMAs that follow accumulate int32 onto that bit
// pattern, so the accumulator read back as a FLOAT is exactly kMagic + x for every |x| <= 127*127*128 < 2^22: the
// int32 -> fp32 conversion (one I2FP per element) disappears into the FFMA that applies the de-quantisation scale,
//   x*c = fma(bits_as_float, c, -kMagic*c),
// which is exactly rounded when kMagic*c is representable, i.e. when c carries at most 22 significant bits
// (magic_scale() rounds the scale to 22 bits: relative change <= 2^-22, below the fp32 rounding of the scale itself).
constexpr float kMagic = 12582912.0f;
constexpr uint32_t kMagicElemA2 = 0x64006400u;       // two fp16 1024.0
constexpr uint32_t kMagicElemB2 = 0x62006200u;       // two fp16 768.0
__device__ __forceinline__ float magic_scale(float c) {   // c rounded to 22 significant bits: kMagic * c is then exact
  return __uint_as_float((__float_as_uint(c) + 2u) & 0xFFFFFFFCu);
}
// Mixed-precision add (PTX ISA 8.6, sm_100: SASS FHADD): float(h.lo / h.hi) + c in one instruction
__device__ __forceinline__ float fhadd_lo(uint32_t h2, float c) {
  float d;
  asm("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; add.rn.f32.f16 %0, lo, %2;}" : "=f"(d) : "r"(h2), "f"(c));
  return d;
}
__device__ __forceinline__ float fhadd_hi(uint32_t h2, float c) {
  float d;
  asm("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; add.rn.f32.f16 %0, hi, %2;}" : "=f"(d) : "r"(h2), "f"(c));
  return d;
}
// four int32 (each already in [-128, 127] or saturated to it) -> four int8 in one word, w0 lowest: 2 x I2IP
__device__ __forceinline__ uint32_t pack_sat_s8x4(int w0, int w1, int w2, int w3) {
  uint32_t t, o;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(t) : "r"(w3), "r"(w2), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(o) : "r"(w1), "r"(w0), "r"(t));
  return o;
}
// two fp32 -> two e4m3 (round to nearest even, saturating), lo in the low byte
__device__ __forceinline__ uint32_t cvt_e4m3x2(float lo, float hi) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 128B-swizzle (Swizzle<3,4,3>): byte offset inside a [rows][128 B] tile whose base is 1024-B aligned.
__device__ __host__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t byte_in_row) {
  return row * 128u + ((((byte_in_row >> 4) ^ (row & 7u)) << 4) | (byte_in_row & 15u));
}


// Launch order of (tile, head) work for causal attention: blockIdx.x = w enumerates groups of G heads; inside a group the
// tile rank r (0 = heaviest) is outermost and the head innermost.  With G = 16 every group starts with its heaviest tiles and
// the last group still ends on its lightest ones (the hardware hands CTAs out in order), while the CTAs running together
// touch at most 16 heads' K / V or Q / dO streams.  G = 1 is plain head-major order.
__device__ __forceinline__ void qa_group_order(int w, int n_heads, int n_tiles, int G, int& rank, int& head) {
  const int per_group = G * n_tiles;
  const int full = n_heads / G;
  int g = w / per_group, hg = G;
  if (g >= full) { g = full; hg = n_heads - full * G; }
  const int rem = w - g * per_group;
  rank = rem / hg;
  head = g * G + (rem - rank * hg);
}
}  // namespace qa
