// TMA tensor-map construction shared by every kernel launcher (and by the self-test library).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>

#include "ttt_internal.h"

namespace tb {

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor [rows][64], box [box_rows][64 cols], 128-B swizzle
int make_token_tmap(CUtensorMap* tm, const void* base, uint64_t rows) { return make_token_tmap_box(tm, base, rows, 64); }
int make_token_tmap_box(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t box_rows) {
  static thread_local char detail[160];
  PFN_encodeTiled enc = get_encode();
  if (!enc) { g_where = "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed"; return -1; }
  cuuint64_t gdim[2] = {64, rows};
  cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(detail, sizeof(detail), "cuTensorMapEncodeTiled(base=%p, rows=%llu) -> CUresult %d", base,
             (unsigned long long)rows, (int)r);
    g_where = detail;
    return -2;
  }
  return 0;
}

}  // namespace tb
