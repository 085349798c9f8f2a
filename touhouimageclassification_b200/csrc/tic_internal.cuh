// Internal launcher prototypes shared between the kernel translation units, api.cu and engine.cu.
#pragma once
#include "tic_common.cuh"

namespace tic {

const char* last_error();

// gemm_tcgen05.cu
int gemm_bf16(const void* A, long long lda, bool a_mn, const void* B, long long ldb, bool b_mn, int M, int N, int K,
              int epilogue, void* out, long long ldo, void* out2, long long ldo2, const float* bias, const void* aux,
              long long ldaux, int aux_int, int splits, cudaStream_t stream);

}  // namespace tic
