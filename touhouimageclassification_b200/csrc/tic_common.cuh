// Shared device/host helpers for the TouhouIC B200 hot path (sm_100a only).
// Inline-PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA/TMEM),
// plus small vector/packing utilities. No torch types anywhere in csrc/.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define TIC_DEVINL __device__ __forceinline__

namespace tic {

// ----------------------------------------------------------------------------------------------
// Error plumbing shared by every launcher (api.cu owns the storage).
// ----------------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);

constexpr int kOk = 0;
constexpr int kErrInvalidArg = 1;
constexpr int kErrCuda = 2;
constexpr int kErrUnsupported = 3;

// ----------------------------------------------------------------------------------------------
// Small math / packing helpers
// ----------------------------------------------------------------------------------------------
TIC_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
TIC_DEVINL float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
TIC_DEVINL float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
TIC_DEVINL float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Exact-erf GELU (transformers.activations.GELUActivation -> torch.nn.functional.gelu, erf form).
TIC_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// d/dx gelu(x) = Phi(x) + x * phi(x)
TIC_DEVINL float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// 2^x on the SFU (MUFU.EX2), flush-to-zero, no range fix-ups: inputs here are <= 0 (softmax / Gaussian exponents).
TIC_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// GELU(x) = x Phi(x) and its derivative Phi(x) + x phi(x) for the GEMM epilogues, whose results are rounded to bf16
// (resolution 2^-9): erfc via Abramowitz-Stegun 7.1.25 (|error| <= 2.5e-5 on erf, i.e. 1.3e-5 on Phi) with one MUFU.RCP
// and one MUFU.EX2 shared by cdf and pdf, constants folded so that the value costs 11 and value + derivative 15 FP32
// instructions (the epilogue of the fc1 GEMM is bound by instruction issue, not by the tensor pipe):
//   h = 0.5 erfc(|x| / sqrt 2) = t (a1' + t (a2' + t a3')) exp(-x^2 / 2),   t = 1 / (1 + p |x| / sqrt 2)
//   gelu(x) = max(x, 0) - |x| h,      gelu'(x) = (x >= 0 ? 1 - h : h) + x exp(-x^2 / 2) / sqrt(2 pi)
TIC_DEVINL void gelu_core(float x, float& h, float& e) {
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(fabsf(x), 0.47047f * 0.70710678118654752440f, 1.0f)));
  e = ex2_approx(x * (x * -0.72134752044448170368f));  // exp(-x^2 / 2)
  float q = fmaf(t, 0.5f * 0.7478556f, 0.5f * -0.0958798f);
  q = fmaf(t, q, 0.5f * 0.3480242f);
  h = (q * t) * e;
}
TIC_DEVINL float gelu_fast(float x) {
  float h, e;
  gelu_core(x, h, e);
  return fmaf(-fabsf(x), h, fmaxf(x, 0.f));
}
TIC_DEVINL float gelu_grad_fast(float x) {
  float h, e;
  gelu_core(x, h, e);
  return fmaf(x * 0.39894228040143267794f, e, x >= 0.f ? 1.0f - h : h);
}
TIC_DEVINL void gelu_and_grad_fast(float x, float& y, float& g) {
  float h, e;
  gelu_core(x, h, e);
  y = fmaf(-fabsf(x), h, fmaxf(x, 0.f));
  g = fmaf(x * 0.39894228040143267794f, e, x >= 0.f ? 1.0f - h : h);
}


// Programmatic dependent launch (see launch_pdl in tic_internal.cuh): let the next kernel of the stream start launching /
// block until the previous kernel has completed and its memory is visible.
TIC_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
TIC_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

TIC_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
TIC_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

TIC_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

TIC_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
TIC_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
TIC_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TIC_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

TIC_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
TIC_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
TIC_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must trap (and surface as a CUDA error) instead of hanging the GPU.
TIC_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
      printf("tic: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
TIC_DEVINL void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
// 2D tiled load global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
TIC_DEVINL void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
TIC_DEVINL void tma_load_3d(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "r"(c2)
      : "memory");
}
// Pull a 3-D box towards L2 without a destination in shared memory: a later tma_load_3d of the same box then pays L2
// latency instead of DRAM latency (used where the shared-memory destination is still occupied when the address is known).
TIC_DEVINL void tma_prefetch_l2_3d(const void* desc, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(
                   reinterpret_cast<uint64_t>(desc)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
TIC_DEVINL void tma_store_2d(const void* desc, const void* smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
TIC_DEVINL void tma_store_3d(const void* desc, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
TIC_DEVINL void tma_store_3d_addr(const void* desc, uint32_t smem_addr, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_addr), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
TIC_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
TIC_DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
TIC_DEVINL void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
TIC_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TIC_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Whole warp. Writes the allocated TMEM base address to *smem_out.
TIC_DEVINL void tmem_alloc(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
TIC_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 inputs, fp32 accumulate). One thread issues.
TIC_DEVINL void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand read from TMEM (bf16 packed), B from smem.
TIC_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed.
TIC_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (one row per thread).
TIC_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
TIC_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
TIC_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
TIC_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
TIC_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

TIC_DEVINL void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two SMs of one cluster execute one 256-row UMMA; only the leader (cluster rank 0)
// issues tcgen05.mma / tcgen05.commit, both CTAs load their own operand halves and own 128 accumulator rows.
// ----------------------------------------------------------------------------------------------
TIC_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
TIC_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `bar` in the LEADER CTA of a pair (clears the peer bit, as CUTLASS' Sm100MmaPeerBitMask).
TIC_DEVINL uint32_t leader_smem_u32(const void* p) { return smem_u32(p) & 0xFEFFFFFFu; }
// arrive on the same-offset mbarrier of CTA `cta` of this cluster
TIC_DEVINL void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remote;\n\t"
      "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [remote];\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
TIC_DEVINL void mbar_arrive_expect_tx_cluster(uint64_t* bar, uint32_t bytes, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remote;\n\t"
      "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
      "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [remote], %2;\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(cta), "r"(bytes)
      : "memory");
}

// ---- cluster launch control (sm_100): a running cluster asks the hardware to cancel a cluster of the same grid that has
// not started yet and takes over its work. The 16-byte response lands asynchronously in shared memory (at the same
// offset in EVERY CTA of the cluster for the multicast form) and completes 16 transaction bytes on the mbarrier.
TIC_DEVINL void clc_try_cancel_multicast(void* response16, uint64_t* bar) {
  asm volatile(
      "clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
      ::"r"(smem_u32(response16)), "r"(smem_u32(bar))
      : "memory");
}
TIC_DEVINL void clc_try_cancel(void* response16, uint64_t* bar) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
               ::"r"(smem_u32(response16)), "r"(smem_u32(bar))
               : "memory");
}
// Decodes a response: returns blockIdx.x of the cancelled cluster's first CTA, or -1 when nothing was left to cancel.
TIC_DEVINL int clc_decode(const void* response16) {
  uint32_t x, ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p1;\n\t"
      ".reg .b128 resp;\n\t"
      "ld.shared.b128 resp, [%2];\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, resp;\n\t"
      "selp.u32 %1, 1, 0, p1;\n\t"
      "mov.u32 %0, 0;\n\t"
      "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, _, _, _}, resp;\n\t"
      "}\n"
      : "=r"(x), "=r"(ok)
      : "r"(smem_u32(response16))
      : "memory");
  return ok ? static_cast<int>(x) : -1;
}

TIC_DEVINL void tma_load_2d_2cta(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0, int32_t c1) {
  // executed by both CTAs of the pair; the transaction bytes are credited to the leader's barrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(leader_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
TIC_DEVINL void tmem_alloc_2cta(uint32_t* smem_out, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_out)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
TIC_DEVINL void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
TIC_DEVINL void umma_bf16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit -> arrive on the same-offset mbarrier in every CTA of `mask`
TIC_DEVINL void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// ----------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout follows the PTX ISA "matrix descriptor" / "instruction descriptor").
// ----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1.
//   [0,14)  start address >> 4        [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4   [46,48) version (1)      [61,64) layout type (2 = SWIZZLE_128B)
TIC_DEVINL uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
//   [4,6) D fmt (1 = f32)  [7,10) A fmt (1 = bf16)  [10,13) B fmt (1 = bf16)
//   [15] A major (1 = MN)  [16] B major (1 = MN)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------
// Host: TMA tensor map encoding through the driver entry point (no link-time libcuda dependency).
// ----------------------------------------------------------------------------------------------
// 2D bf16 tensor map, 128B swizzle: dim0 = contiguous dimension (elements), dim1 = rows,
// row pitch in elements, box = box0 x box1 elements (box0 * 2 bytes must be <= 128).
int encode_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t pitch_elems,
                        uint32_t box0, uint32_t box1);

int encode_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                        uint64_t pitch1_elems, uint64_t pitch2_elems, uint32_t box0, uint32_t box1);

int encode_tmap_3d_bf16_sw(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                           uint64_t pitch1_elems, uint64_t pitch2_elems, uint32_t box0, uint32_t box1, int swizzle_bytes);

}  // namespace tic
