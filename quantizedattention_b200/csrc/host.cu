// C-ABI plumbing: version, thread-local last-error, launch checks, TMA descriptor encoding through the driver
// entry point (no link-time dependency on libcuda).
#include "qa_host.h"
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

int qa_make_tmap(CUtensorMap* out, const void* gptr, CUtensorMapDataType elem, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, int swizzle) {
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
  return QA_OK;
}
