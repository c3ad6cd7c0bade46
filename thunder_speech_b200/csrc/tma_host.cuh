// Host-side CUtensorMap construction.  cuTensorMapEncodeTiled is fetched through
// cudaGetDriverEntryPoint so that the library has no link-time dependency on libcuda.so (it must load on
// the CPU-only build container for the symbol check).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "ts_common.cuh"

namespace ts {
int option_tma_l2();
namespace tma {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline int encode(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  TS_REQUIRE(fn != nullptr, TS_ERR_NO_DEVICE, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  TS_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, TS_ERR_INVALID, "TMA base pointer must be 16-byte aligned");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  (CUtensorMapL2promotion)option_tma_l2(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TS_REQUIRE(r == CUDA_SUCCESS, TS_ERR_INVALID,
             "cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu strides %llu %llu box %u %u %u", (int)r,
             rank, (unsigned long long)dims[0], (unsigned long long)dims[1], rank > 2 ? (unsigned long long)dims[2] : 0ull,
             (unsigned long long)strides_bytes[0], rank > 2 ? (unsigned long long)strides_bytes[1] : 0ull, box[0], box[1],
             rank > 2 ? box[2] : 0u);
  return TS_OK;
}

// [d1, d0] bf16, d0 contiguous; stride1 = bytes between consecutive d1 indices
inline int make_2d_bf16(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t stride1, uint32_t box0,
                        uint32_t box1) {
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {stride1};
  cuuint32_t box[2] = {box0, box1};
  return encode(m, ptr, 2, dims, strides, box);
}

// [d2, d1, d0] bf16, d0 contiguous
inline int make_3d_bf16(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1,
                        uint64_t stride2, uint32_t box0, uint32_t box1, uint32_t box2) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1, stride2};
  cuuint32_t box[3] = {box0, box1, box2};
  return encode(m, ptr, 3, dims, strides, box);
}

}  // namespace tma
}  // namespace ts
