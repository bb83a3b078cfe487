// Host-side construction of TMA tensor maps (CUtensorMap) without linking libcuda:
// cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

namespace lcbi {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// Text describing the last failed encode (per thread): which view was rejected and why; capi.cu appends it to
// lcbi_last_error() so that a rejected tensor map names its pointer, dims, strides and box.
inline char* tmap_error_text() {
  static thread_local char buf[384] = "";
  return buf;
}

// rank-R tiled map. dims[0] is the contiguous dimension; strides_bytes[i] is the byte stride of
// dims[i+1] (R-1 entries, multiples of 16). Returns 0 on success.
inline int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base,
                     const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                     CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return -1;
  cuuint64_t gdim[5], gstr[5];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT) {
    // A driver-API call needs a context bound to THIS thread; a fresh thread (e.g. autograd's backward worker whose
    // first node is one of our kernels) has none until its first runtime call. Bind the primary context and retry.
    cudaFree(nullptr);
    r = enc(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    char* t = tmap_error_text();
    int n = std::snprintf(t, 384, "cuTensorMapEncodeTiled -> %d: base %p rank %d dtype %d swizzle %d dims", (int)r, base, rank,
                          (int)dt, (int)swz);
    for (int i = 0; i < rank && n < 360; ++i) n += std::snprintf(t + n, 384 - n, " %llu", (unsigned long long)gdim[i]);
    n += std::snprintf(t + n, 384 - n, " strides(B)");
    for (int i = 0; i + 1 < rank && n < 360; ++i) n += std::snprintf(t + n, 384 - n, " %llu", (unsigned long long)gstr[i]);
    n += std::snprintf(t + n, 384 - n, " box");
    for (int i = 0; i < rank && n < 370; ++i) n += std::snprintf(t + n, 384 - n, " %u", bx[i]);
  }
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace lcbi
