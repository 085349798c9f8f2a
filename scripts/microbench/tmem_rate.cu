// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM with 1..16 warps issuing back to back.
#include "../../touhouimageclassification_b200/csrc/tic_common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace tic;

template <int OP>  // 0: ld x32, 1: st x16, 2: ld x16
__global__ void __launch_bounds__(512, 1) k(int nwarps, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  const int warp = threadIdx.x >> 5;
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    const uint32_t lane_addr = tb + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t r[32], w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = i + threadIdx.x;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (OP == 0) {
        tmem_ld_32x32b_x32(lane_addr + (i & 1) * 32, r);
        tmem_ld_wait();
        acc += r[0] + r[31];
      } else if (OP == 2) {
        uint32_t q[16];
        tmem_ld_32x32b_x16(lane_addr + (i & 3) * 16, q);
        tmem_ld_wait();
        acc += q[0] + q[15];
      } else {
        tmem_st_32x32b_x16(lane_addr + (i & 3) * 16, w);
        tmem_st_wait();
      }
    }
    t1 = clock64();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* out;
  uint32_t* sink;
  cudaMallocManaged(&out, 16);
  cudaMalloc(&sink, 16);
  const int iters = 2048;
  for (int op = 0; op < 3; ++op)
    for (int nw : {1, 2, 4, 8, 16}) {
      if (op == 0) k<0><<<148, 512>>>(nw, iters, out, sink);
      else if (op == 1) k<1><<<148, 512>>>(nw, iters, out, sink);
      else k<2><<<148, 512>>>(nw, iters, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      const double bytes = (op == 0 ? 4096.0 : 2048.0) * nw * iters;
      printf("%s warps=%2d: %.1f clk per op per warp, %.1f B/clk/SM\n", op == 0 ? "ld.x32" : op == 1 ? "st.x16" : "ld.x16", nw,
             (double)out[0] / iters, bytes / out[0]);
    }
  return 0;
}
