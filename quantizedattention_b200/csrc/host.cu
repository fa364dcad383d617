// C-ABI plumbing: version, thread-local last-error, launch checks, TMA descriptor encoding through the driver
// entry point (no link-time dependency on libcuda).
#include "qa_host.h"

extern "C" size_t qa_k_mean_workspace_bytes(int B, int H, int S, int D);
#include <stdio.h>
#include <string.h>
#include <mutex>

static thread_local char g_err[512] = "";

int qa_fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int qa_check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return QA_ERR_CUDA;
  }
  return QA_OK;
}

extern "C" const char* qa_last_error(void) { return g_err; }
extern "C" int qa_version(void) { return 100; }   // 0.1.0

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// Per-process cache of encoded tensor maps, keyed by everything that goes into cuTensorMapEncodeTiled: a training step
// re-launches the same kernels over the same buffers, and each launch needs 3-5 maps.  Direct-mapped, mutex-guarded (the
// autograd engine calls backward from its own thread); a hit is a 128-byte copy.
namespace {
struct TmapKey {
  const void* ptr;
  int elem, rank, swizzle;
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapSlot {
  bool valid = false;
  TmapKey key;
  CUtensorMap map;
};
constexpr int kTmapSlots = 256;
TmapSlot g_tmap_cache[kTmapSlots];
std::mutex g_tmap_mutex;
uint64_t tmap_hash(const TmapKey& k) {
  const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < sizeof(TmapKey); ++i) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}
}  // namespace

int qa_make_tmap(CUtensorMap* out, const void* gptr, CUtensorMapDataType elem, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
  if (rank < 1 || rank > 5) return qa_fail(QA_ERR_SHAPE, "qa_make_tmap: rank must be 1..5");
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = gptr; key.elem = (int)elem; key.rank = rank; key.swizzle = swizzle;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; if (i > 0) key.strides[i - 1] = strides_bytes[i - 1]; }
  TmapSlot& slot = g_tmap_cache[tmap_hash(key) % kTmapSlots];
  {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    if (slot.valid && slot.key == key) { *out = slot.map; return QA_OK; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return qa_fail(QA_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = swizzle == 3   ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B
                                         : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, elem, (cuuint32_t)rank, const_cast<void*>(gptr), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return qa_fail(QA_ERR_DRIVER, buf);
  }
  {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    slot.key = key; slot.map = *out; slot.valid = true;
  }
  return QA_OK;
}

// Caller-owned workspaces (the library never allocates): bytes each op needs, so that a host in another language can
// size its buffers without knowing the kernels.  op: QA_WS_* of include/qattn.h.
extern "C" size_t qa_workspace_bytes(int op, int B, int H, int S, int D) {
  const size_t BH = (size_t)B * (size_t)H, N = BH * (size_t)S;
  switch (op) {
    case 0: return qa_k_mean_workspace_bytes(B, H, S, D);      // QA_WS_K_MEAN
    case 1: return N * (size_t)D * sizeof(float);               // QA_WS_INT8_BWD_DQ: fp32 dQ accumulator, zero-initialised
    case 2: return N * sizeof(float);                           // QA_WS_INT8_BWD_ROWSUM: fp32 rowsum(dS), zero-initialised
    case 3: return 6 * N * (size_t)D * 2;                       // QA_WS_JVP_BF16_OPERANDS: six bf16 copies of the fp32 inputs
    default: return 0;
  }
}
