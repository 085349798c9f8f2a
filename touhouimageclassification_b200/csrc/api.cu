// extern "C" surface of libtic_b200.so (declared in include/tic_b200.h). Thin argument adapters only.
#include "tic_b200.h"
#include "tic_internal.cuh"

using namespace tic;

extern "C" {

TIC_API int tic_abi_version(void) { return 1; }
TIC_API const char* tic_last_error(void) { return last_error(); }

TIC_API int tic_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                          int M, int N, int K, int epilogue, void* out, int64_t ldo, void* out2, int64_t ldo2,
                          const float* bias, const void* aux, int64_t ldaux, int aux_int, int splits, void* stream) {
  return gemm_bf16(A, lda, a_mn_major != 0, B, ldb, b_mn_major != 0, M, N, K, epilogue, out, ldo, out2, ldo2, bias,
                   aux, ldaux, aux_int, splits, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
