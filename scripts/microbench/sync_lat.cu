// Microbenchmark: synchronisation latencies inside one CTA (cycles, clock64 on one SM)
//  (1) mbarrier.arrive by warp A -> try_wait success in spinning warp B
//  (2) n x tcgen05.mma (M=128,N=64,K=16, SS) + tcgen05.commit by warp A -> try_wait success in warp B and in warp A
//  (3) ping-pong round trip: A arrives on bar1, B waits bar1 then arrives on bar2, A waits bar2 (per round)
#include "../../touhouimageclassification_b200/csrc/tic_common.cuh"
#include <cstdio>
using namespace tic;

__device__ __forceinline__ void spin_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }

__global__ void __launch_bounds__(64, 1) k(long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ long long tstamp[8];
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
    __syncwarp();
    tmem_alloc(&slot, 512);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  const int warp = threadIdx.x >> 5;
  const bool leader = elect_one();
  const uint32_t idesc = make_idesc_bf16(128, 64, false, false);
  const uint64_t da = make_smem_desc_sw128(smem_u32(smem), 0, 1024), db = make_smem_desc_sw128(smem_u32(smem + 32768), 0, 1024);
  // ---- (1) arrive -> wake
  for (int rep = 0; rep < 4; ++rep) {
    __syncthreads();
    if (warp == 0 && leader) { for (volatile int d = 0; d < 2000; ++d) {} tstamp[0] = clock64(); mbar_arrive(&bar[0]); }
    if (warp == 1 && leader) { spin_wait(&bar[0], rep & 1); tstamp[1] = clock64(); }
    __syncthreads();
    if (threadIdx.x == 0) out[rep] = tstamp[1] - tstamp[0];
  }
  // ---- (2) n MMAs + commit -> visible
  const int ns[4] = {1, 4, 8, 16};
  for (int c = 0; c < 4; ++c) {
    for (int rep = 0; rep < 2; ++rep) {
      const int ph = (c * 2 + rep) & 1;
      __syncthreads();
      if (warp == 0 && leader) {
        for (volatile int d = 0; d < 2000; ++d) {}
        tstamp[0] = clock64();
        for (int i = 0; i < ns[c]; ++i) umma_bf16_ss(tb + 256, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1u);
        umma_commit(&bar[1]);
        tstamp[2] = clock64();
        spin_wait(&bar[1], ph);
        tstamp[3] = clock64();
      }
      if (warp == 1 && leader) { spin_wait(&bar[1], ph); tstamp[1] = clock64(); }
      __syncthreads();
      if (threadIdx.x == 0 && rep == 1) { out[8 + c * 3] = tstamp[1] - tstamp[0]; out[9 + c * 3] = tstamp[3] - tstamp[0]; out[10 + c * 3] = tstamp[2] - tstamp[0]; }
    }
  }
  // ---- (3) ping-pong
  __syncthreads();
  const int rounds = 256;
  long long t0 = clock64();
  if (warp == 0 && leader) for (int i = 0; i < rounds; ++i) { mbar_arrive(&bar[2]); spin_wait(&bar[3], i & 1); }
  if (warp == 1 && leader) for (int i = 0; i < rounds; ++i) { spin_wait(&bar[2], i & 1); mbar_arrive(&bar[3]); }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[30] = (t1 - t0) / rounds;
  // ---- (4) ping-pong with the library mbar_wait (try_wait fast path + timeout loop)
  __syncthreads();
  if (threadIdx.x == 0) { mbar_init(&bar[2], 1); mbar_init(&bar[3], 1); fence_mbar_init(); }
  __syncthreads();
  t0 = clock64();
  if (warp == 0 && leader) for (int i = 0; i < rounds; ++i) { mbar_arrive(&bar[2]); mbar_wait(&bar[3], i & 1); }
  if (warp == 1 && leader) for (int i = 0; i < rounds; ++i) { mbar_wait(&bar[2], i & 1); mbar_arrive(&bar[3]); }
  t1 = clock64();
  if (threadIdx.x == 0) out[31] = (t1 - t0) / rounds;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<<<1, 64, 100 * 1024>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  printf("arrive -> wake (other warp): %lld %lld %lld %lld clk\n", out[0], out[1], out[2], out[3]);
  const int ns[4] = {1, 4, 8, 16};
  for (int c = 0; c < 4; ++c)
    printf("%2d x MMA(128x64x16 SS) + commit: issue done %lld, visible other warp %lld, same warp %lld clk (exec floor %d)\n", ns[c],
           out[10 + c * 3], out[8 + c * 3], out[9 + c * 3], ns[c] * 48);
  printf("ping-pong round trip (2 hops), raw spin: %lld clk; with mbar_wait helper: %lld clk\n", out[30], out[31]);
  return 0;
}
