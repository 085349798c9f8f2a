// Host-side runtime shared by every launcher: last-error storage, launch checking and TMA tensor-map
// encoding through the driver entry point (so the library has no link-time dependency on libcuda).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "tic_common.cuh"

namespace tic {

namespace {
thread_local char g_last_error[512] = "";

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;
std::once_flag g_encode_once;

void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
}
}  // namespace

const char* last_error() { return g_last_error; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(kErrCuda, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return kOk;
}

int encode_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t pitch_elems,
                        uint32_t box0, uint32_t box1) {
  std::call_once(g_encode_once, load_encode);
  if (g_encode == nullptr) return set_error(kErrCuda, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(kErrCuda, "cuTensorMapEncodeTiled failed (%d) base=%p dims=%llu,%llu pitch=%llu box=%u,%u", (int)r,
                     base, (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)pitch_elems, box0,
                     box1);
  return kOk;
}

}  // namespace tic
