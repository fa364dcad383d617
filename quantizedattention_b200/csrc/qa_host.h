// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

enum {
  QA_OK = 0,
  QA_ERR_SHAPE = -1,
  QA_ERR_ALIGN = -2,
  QA_ERR_WORKSPACE = -3,
  QA_ERR_CUDA = -4,
  QA_ERR_ARCH = -5,
  QA_ERR_DRIVER = -6,
};

enum { QA_FLAG_NEAREST = 1, QA_FLAG_CAUSAL = 2, QA_FLAG_BWD_8WARP = 4 };   // `flags` of qa_int8_fwd / qa_int8_bwd (include/qattn.h)

int qa_fail(int code, const char* msg);            // records msg in the thread-local error slot, returns code
int qa_check_launch(const char* where);            // cudaGetLastError -> QA_OK / QA_ERR_CUDA (never synchronises)

// 2-D / 3-D tiled tensor map over a row-major tensor.  dims/box innermost first; strides in BYTES for dims 1..rank-1.
// swizzle: 0 none, 1 = 32B, 2 = 64B, 3 = 128B.  elem: CU_TENSOR_MAP_DATA_TYPE_*.
int qa_make_tmap(CUtensorMap* out, const void* gptr, CUtensorMapDataType elem, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box, int swizzle);
