// Microbenchmark: issue rates per SM of the instructions the softmax / GELU passes are made of -- FFMA (3 registers),
// FFMA2 (fma.rn.f32x2), MUFU.EX2, cvt.rn.bf16x2.f32 -- with 4 / 8 / 16 warps per SM, 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned long long pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int OP>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out, float* sink, float c) {
  float a[8];
  unsigned long long A[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; A[i] = pk(a[i], a[i] + 0.5f); }
  const unsigned long long C = pk(c, c), Dd = pk(0.25f, 0.75f);
  const float d = 0.25f * c;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A[i]) : "l"(C), "l"(Dd));
      if (OP == 2) a[i] = ex2(a[i]);
      if (OP == 3) {  // one ex2 + 3 fma (the softmax inner body, scalar)
        a[i] = ex2(a[i]);
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
      }
      if (OP == 4) {
        uint32_t w;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(a[i]), "f"(a[(i + 1) & 7]));
        a[i] = __uint_as_float(w);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(static_cast<uint32_t>(A[i]));
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* out;
  float* sink;
  cudaMallocManaged(&out, 16);
  cudaMalloc(&sink, 16);
  const int iters = 4096;
  const char* names[5] = {"FFMA", "FFMA2", "MUFU.EX2", "EX2+3FFMA", "F2F.BF16X2"};
  for (int op = 0; op < 5; ++op)
    for (int nw : {4, 8, 16}) {
      if (op == 0) k<0><<<148, nw * 32>>>(iters, out, sink, 1.0001f);
      if (op == 1) k<1><<<148, nw * 32>>>(iters, out, sink, 1.0001f);
      if (op == 2) k<2><<<148, nw * 32>>>(iters, out, sink, 1.0001f);
      if (op == 3) k<3><<<148, nw * 32>>>(iters, out, sink, 1.0001f);
      if (op == 4) k<4><<<148, nw * 32>>>(iters, out, sink, 1.0001f);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      const double instr = 8.0 * iters * nw * (op == 3 ? 4 : 1);
      printf("%-10s warps=%2d: %.2f warp-instr/clk/SM (%.2f clk per warp-instr per SMSP)\n", names[op], nw, instr / out[0],
             out[0] / (instr / 4));
    }
  return 0;
}
