// Host helper: cuTensorMapEncodeTiled through the runtime's driver-entry-point query (the library does not link
// against libcuda).  Used by the kernels that stage operands with tensor-map TMA copies.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace avc {

typedef CUresult (*PFN_tmap_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_tmap_encode tmap_encode_fn() {
  static PFN_tmap_encode fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmap_encode)f;
  }
  return fn;
}

// A4 activation [B][C/4][T][4] (sample stride bstride floats) as a 4-D fp32 tensor (4 floats, chunk, time, sample) with a
// box of (4, box_chunks, box_rows, box_samples); CU_TENSOR_MAP_SWIZZLE_* as given.  Returns CUDA_SUCCESS or the error.
static inline CUresult tmap_a4_chunk_time_sample(CUtensorMap* tm, const float* base, int C, int T, int B, long long bstride, int box_chunks,
                                                  int box_rows, int box_samples, CUtensorMapSwizzle swz) {
  PFN_tmap_encode enc = tmap_encode_fn();
  if (!enc) return CUDA_ERROR_NOT_SUPPORTED;
  const cuuint64_t gdim[4] = {4, (cuuint64_t)(C / 4), (cuuint64_t)T, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)T * 16u, 16, (cuuint64_t)bstride * 4u};
  const cuuint32_t box[4] = {4, (cuuint32_t)box_chunks, (cuuint32_t)box_rows, (cuuint32_t)box_samples};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace avc
