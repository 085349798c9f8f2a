// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, cta_group::1) for the small shapes the attention kernels use,
// with a cheap issue loop (descriptors precomputed, 8 MMAs unrolled per iteration, accumulate flag immediate).
// Two issue styles: STYLE 0 = `if (lane == 0)` divergent single thread; STYLE 1 = whole warp + elect_one.
#include "../../touhouimageclassification_b200/csrc/tic_common.cuh"
#include <cstdio>
#include <cstdlib>
using namespace tic;

TIC_DEVINL void mma_ss_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
TIC_DEVINL void mma_ts_acc(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc) : "memory");
}

template <int MODE, int STYLE>
__global__ void __launch_bounds__(128, 1) k(int N, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(&slot, 512);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (threadIdx.x < 32) {
    const bool leader = STYLE == 0 ? (threadIdx.x == 0) : elect_one();
    const uint32_t aA = smem_u32(smem), aB = smem_u32(smem + 64 * 1024);
    const uint32_t idesc = MODE == 0 ? make_idesc_bf16(128, N, false, false)
                         : MODE == 1 ? make_idesc_bf16(128, N, false, true) : make_idesc_bf16(128, N, true, true);
    const uint64_t da = MODE == 2 ? make_smem_desc_sw128(aA, 16384, 1024) : make_smem_desc_sw128(aA, 0, 1024);
    const uint64_t db = MODE == 0 ? make_smem_desc_sw128(aB, 0, 1024) : make_smem_desc_sw128(aB, 8192, 1024);
    const uint32_t d0 = tb + 256;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (leader) {
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (MODE == 0) mma_ss_acc(d0, da + 2 * (u & 3), db + 2 * (u & 3), idesc);
          else if (MODE == 1) mma_ts_acc(d0, tb + 8 * (u & 3), db + 128 * (u & 3), idesc);
          else mma_ss_acc(d0, da + 128 * (u & 3), db + 128 * (u & 3), idesc);
        }
      }
      t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int MODE, int STYLE>
void run(const char* name, long long* out) {
  cudaFuncSetAttribute(k<MODE, STYLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int N : {16, 64, 128, 256}) {
    const int iters = 1024;
    k<MODE, STYLE><<<148, 128, 200 * 1024>>>(N, iters, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    printf("%s style=%d M=128 N=%3d K=16: issue %.1f clk/mma, complete %.1f clk/mma (floor %.0f)\n", name, STYLE, N,
           (double)out[0] / iters, (double)out[1] / iters, 128.0 * N / 256);
  }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  run<0, 0>("SS  K/K ", out); run<0, 1>("SS  K/K ", out);
  run<1, 0>("TS  -/MN", out); run<1, 1>("TS  -/MN", out);
  run<2, 0>("SS MN/MN", out); run<2, 1>("SS MN/MN", out);
  return 0;
}
