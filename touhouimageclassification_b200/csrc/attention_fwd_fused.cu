// Fused multi-head attention FORWARD for sequences of up to 224 tokens (the 197-token ViT-*/16 224x224 case) as a
// persistent, warp-specialised kernel. Replaces F.scaled_dot_product_attention (modeling_vit.py:232-246 [a6]),
// non-causal, no mask, dropout 0, head_dim 64.
//
// One CTA per SM walks the (image, head) items. For each item all keys fit in ONE block, so there is no online
// rescaling: S = Q K^T (SS MMA, M = 128 queries, N = keys rounded up to 16), row max, P = exp2(S c - max), O = P V
// (TS MMA, P read from TMEM as packed bf16 over the S columns it came from), O / rowsum -> bf16.
//   warps 0-3   softmax group of query tile 0 (rows 0-127), one query row per thread (TMEM lane == row)
//   warps 4-7   softmax group of query tile 1 (rows 128-255)
//   warp  8     one elected thread: TMA loads (Q, K, V of the NEXT item are prefetched into the other operand stage
//               while this item is computed) and all MMAs
// The two query tiles run side by side (two softmax warps per SM sub-partition keep its MUFU and FMA pipes busier than
// one does: staggering the tiles with named barriers measured slower). Each tile owns 256 TMEM columns:
// S fp32 [0, 224) -> P bf16 [0, 112), O fp32 [128, 192) (aliases the dead tail of S).
// Outputs leave through per-warp shared-memory tiles and TMA stores (rows past N are clipped by the tensor map):
// direct 16-byte stores at a 2 KB row pitch cost ~300 clk per instruction.
#include "tic_internal.cuh"


namespace tic {
namespace {

constexpr int FF_THREADS = 288;
constexpr int FF_HD = 64;
constexpr int FF_KV = 224;                                  // key rows staged per item
constexpr int FF_Q_BYTES = 256 * 128;                       // 32 KB
constexpr int FF_KV_BYTES = FF_KV * 128;                    // 28 KB
constexpr int FF_STAGE_BYTES = FF_Q_BYTES + 2 * FF_KV_BYTES;  // 88 KB
constexpr int FF_OUT_BYTES = 8 * 4096;                      // per softmax warp: 32 rows x 128 B
constexpr int FF_SMEM_USED = 2 * FF_STAGE_BYTES + FF_OUT_BYTES + 256;
constexpr int FF_SMEM = FF_SMEM_USED + 1024;
constexpr uint32_t FF_O_COL = 128;
constexpr float FF_LOG2E = 1.4426950408889634f;
constexpr float FF_LN2 = 0.6931471805599453f;

TIC_DEVINL void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
TIC_DEVINL void ff_st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(FF_THREADS, 1)
attn_fwd_fused_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o,
                      float* __restrict__ lse, int N, int Nq, int H, int num_items, float scale,
                      long long* __restrict__ trace) {
  // N = keys per item; Nq = queries per item (the first Nq tokens of each image: Nq = N normally, Nq = 1 when only the
  // CLS row of the last encoder layer is needed)
  // trace (dev tool, normally NULL): clock64 stamps of CTA 0, third item -- [0..31] warp 0, [32..63] warp 4, [64..] MMA thread
#ifdef TIC_ATTN_TRACE  // development build only (-DTIC_ATTN_TRACE): the shipped library has no tracing code
#define FF_STAMP(slot) do { if (trace != nullptr && blockIdx.x == 0 && it == 2) trace[slot] = clock64(); } while (0)
#else
#define FF_STAMP(slot) do { } while (0)
#endif
  extern __shared__ uint8_t ff_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ff_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sOut = smem + 2 * FF_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + FF_OUT_BYTES);
  uint64_t* bar_ld = bars + 0;     // [2] operand stage loaded
  uint64_t* bar_s = bars + 2;      // [2] S of a query tile is in TMEM
  uint64_t* bar_p = bars + 4;      // [2] P of a query tile written (4 warp arrivals)
  uint64_t* bar_o = bars + 6;      // [2] P V of a query tile has completed
  uint64_t* bar_ofree = bars + 8;  // [2] the tile's warps have read O out of TMEM (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (Nq + 127) >> 7;         // query tiles in use (1 or 2)
  const int nk = (N + 15) & ~15;           // keys rounded up to the UMMA N / K granularity (padded keys are zero rows)

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_ld[i], 1);
        mbar_init(&bar_s[i], 1);
        mbar_init(&bar_p[i], 4);
        mbar_init(&bar_o[i], 1);
        mbar_init(&bar_ofree[i], 4);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (elect_one()) {
      // ------------------------------------------------------------------------------ TMA + MMA issue thread
      const uint32_t idesc_s = make_idesc_bf16(128, nk, false, false);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, FF_HD, false, true);
      const int ksteps = nk >> 4;
      auto issue_loads = [&](int item, int stage) {
        const int h = item % H, b = item / H;
        uint8_t* st = smem + stage * FF_STAGE_BYTES;
        mbar_arrive_expect_tx(&bar_ld[stage], FF_STAGE_BYTES);
        tma_load_3d(st + FF_Q_BYTES, &tm_k, &bar_ld[stage], h * FF_HD, 0, b);
        tma_load_3d(st, &tm_q, &bar_ld[stage], h * FF_HD, 0, b);
        tma_load_3d(st + FF_Q_BYTES + FF_KV_BYTES, &tm_v, &bar_ld[stage], h * FF_HD, 0, b);
      };
      issue_loads(blockIdx.x, 0);
      for (int it = 0, item = blockIdx.x; item < num_items; ++it, item += gridDim.x) {
        const int stage = it & 1;
        const uint32_t ph = it & 1;  // every per-tile barrier completes exactly once per item
        const uint32_t aQ = smem_u32(smem + stage * FF_STAGE_BYTES), aK = aQ + FF_Q_BYTES, aV = aK + FF_KV_BYTES;
        mbar_wait(&bar_ld[stage], (it >> 1) & 1);
        tc_fence_after();
        const uint64_t dk = make_smem_desc_sw128(aK, 0, 1024);
        for (int t = 0; t < nqt; ++t) {
          if (it > 0) {  // the previous item's O of this tile has left TMEM
            mbar_wait(&bar_ofree[t], ph ^ 1);
            tc_fence_after();
          }
          const uint64_t dq = make_smem_desc_sw128(aQ + t * 16384, 0, 1024);
#pragma unroll
          for (int k = 0; k < FF_HD / 16; ++k) umma_bf16_ss(tmem_base + t * 256, dq + 2 * k, dk + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          umma_commit(&bar_s[t]);
        }
        // The other operand stage was last read by the previous item's P V products, which have completed (their O has
        // even been read back): prefetch the next item into it.
        if (item + static_cast<int>(gridDim.x) < num_items) issue_loads(item + gridDim.x, stage ^ 1);
        const uint64_t dv = make_smem_desc_sw128(aV, 8192, 1024);
        for (int t = 0; t < nqt; ++t) {
          mbar_wait(&bar_p[t], ph);
          tc_fence_after();
          for (int k = 0; k < ksteps; ++k)
            umma_bf16_ts(tmem_base + t * 256 + FF_O_COL, tmem_base + t * 256 + 8 * k, dv + 128 * k, idesc_o, k > 0 ? 1u : 0u);
          umma_commit(&bar_o[t]);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------- softmax warps
    const int t = warp >> 2, quad = warp & 3;  // query tile, TMEM lane quadrant
    const uint32_t lane_addr = tmem_base + t * 256 + (static_cast<uint32_t>(quad * 32) << 16);
    const float c2 = scale * FF_LOG2E;
    const int row = t * 128 + quad * 32 + lane;      // query row within the item
    const bool warp_active = t * 128 + quad * 32 < Nq;  // warps whose 32 rows are all padding only keep the barriers moving
    const uint32_t out_tile = smem_u32(sOut) + warp * 4096;
    const int nchunk = (N + 31) >> 5;                // 32-column chunks that hold at least one valid key
    const int tail = N - (nchunk - 1) * 32;          // valid keys in the last chunk (1..32)
    if (t < nqt) {
      for (int it = 0, item = blockIdx.x; item < num_items; ++it, item += gridDim.x) {
        const int h = item % H, b = item / H;
        const uint32_t ph = it & 1;
#define FF_WSTAMP(slot) do { if (lane == 0 && quad == 0) FF_STAMP(t * 32 + (slot)); } while (0)
        FF_WSTAMP(0);
        mbar_wait(&bar_s[t], ph);
        tc_fence_after();
        FF_WSTAMP(1);
        float mx = 0.f, l = 1.f;
        if (warp_active) {
          // pass 1: row maximum of the raw scores over the valid keys
          float raw_max = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < nchunk; ++c) {
            const int nvalid = c == nchunk - 1 ? tail : 32;
            if (nvalid > 8) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(lane_addr + c * 32, r);
              tmem_ld_wait();
              if (nvalid == 32) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) raw_max = fmaxf(raw_max, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < nvalid) raw_max = fmaxf(raw_max, __uint_as_float(r[i]));
              }
            } else {
              uint32_t r[8];
              tmem_ld_32x32b_x8(lane_addr + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (i < nvalid) raw_max = fmaxf(raw_max, __uint_as_float(r[i]));
            }
          }
          mx = raw_max * c2;
          FF_WSTAMP(2);
          // pass 2: p = exp2(s * c2 - mx), row sum, packed bf16 P over the S columns (zeros past the last key).
          // (Keeping the next chunk's TMEM load in flight and splitting the accumulators measured slower: the extra
          // register copies cost more than the exposed load latency with two softmax warps per sub-partition.)
          float sum = 0.f;
          const float neg_mx = -mx;
#pragma unroll 1
          for (int c = 0; c < (nk + 31) >> 5; ++c) {
            const int nvalid = c < nchunk - 1 ? 32 : (c == nchunk - 1 ? tail : 0);
            uint32_t w[16];
            if (nvalid > 8) {
              uint32_t r[32];
              tmem_ld_32x32b_x32(lane_addr + c * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), c2, neg_mx));
                float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c2, neg_mx));
                if (nvalid < 32) {
                  if (2 * i >= nvalid) p0 = 0.f;
                  if (2 * i + 1 >= nvalid) p1 = 0.f;
                }
                sum += p0 + p1;
                w[i] = pack_bf16x2(p0, p1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) w[i] = 0u;
              if (nvalid > 0) {
                uint32_t r[8];
                tmem_ld_32x32b_x8(lane_addr + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * i]), c2, neg_mx));
                  float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * i + 1]), c2, neg_mx));
                  if (2 * i >= nvalid) p0 = 0.f;
                  if (2 * i + 1 >= nvalid) p1 = 0.f;
                  sum += p0 + p1;
                  w[i] = pack_bf16x2(p0, p1);
                }
              }
            }
            tmem_st_32x32b_x16(lane_addr + c * 16, w);
          }
          tmem_st_wait();
          l = sum;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_p[t]);
        FF_WSTAMP(3);

        // ---- O / l -> bf16 through this warp's staging tile and a TMA store; logsumexp for the backward pass
        mbar_wait(&bar_o[t], ph);
        tc_fence_after();
        FF_WSTAMP(4);
        uint32_t packed[32];
        if (warp_active) {
          const float inv_l = 1.0f / l;
#pragma unroll
          for (int c = 0; c < FF_HD / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(lane_addr + FF_O_COL + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              packed[c * 16 + i] = pack_bf16x2(__uint_as_float(r[2 * i]) * inv_l, __uint_as_float(r[2 * i + 1]) * inv_l);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_ofree[t]);  // the tile's TMEM region may be overwritten by the next item's S
        FF_WSTAMP(5);
        if (warp_active) {
          if (lane == 0) tma_store_wait_read<0>();  // the previous item's store has finished reading the staging tile
          __syncwarp();
          const uint32_t base = out_tile + lane * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            ff_st_shared_v4(base + ((i ^ (lane & 7)) << 4), packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d_addr(&tm_o, out_tile, h * FF_HD, t * 128 + quad * 32, b);
            tma_store_commit();
          }
          if (lse != nullptr && row < Nq) lse[static_cast<long long>(item) * Nq + row] = (mx + log2f(l)) * FF_LN2;
        }
        FF_WSTAMP(6);
      }
      if (lane == 0) tma_store_wait_read<0>();  // the staging tile must outlive the last TMA store
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 8) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// q/k/v: [B*N, ...] with row pitch ld, head h at column h*64; o: [B*N, H*64] pitch ldo; lse: [B, H, Nq] or NULL.
// Only the first Nq tokens of every image act as queries (their rows of o / lse are written); all N are keys.
int attention_fwd_fused(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, float* lse, int B,
                        int N, int Nq, int H, float scale, cudaStream_t stream) {
  if (N > FF_KV) return set_error(kErrUnsupported, "attention_fwd_fused: N=%d > %d", N, FF_KV);
  if (Nq <= 0 || Nq > N) return set_error(kErrInvalidArg, "attention_fwd_fused: Nq=%d must be in [1, N=%d]", Nq, N);
  CUtensorMap tq, tk, tv, to;
  const uint64_t D = static_cast<uint64_t>(H) * FF_HD;
  int rc = encode_tmap_3d_bf16(&tq, q, D, Nq, B, ld, static_cast<uint64_t>(N) * ld, 64, 256);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tk, k, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, FF_KV);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&tv, v, D, N, B, ld, static_cast<uint64_t>(N) * ld, 64, FF_KV);
  if (rc) return rc;
  rc = encode_tmap_3d_bf16(&to, o, D, Nq, B, ldo, static_cast<uint64_t>(N) * ldo, 64, 32);  // one warp's 32-row tile
  if (rc) return rc;
  if (int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(attn_fwd_fused_kernel), FF_SMEM, "attention_fwd_fused")) return rc2;
  const int num_sms = device_sm_count();
  const int items = B * H;
  dim3 grid(items < num_sms ? items : num_sms);
  long long* trace = nullptr;
#ifdef TIC_ATTN_TRACE
  cudaMallocManaged(&trace, 128 * sizeof(long long));
  for (int i = 0; i < 128; ++i) trace[i] = 0;
#endif
  attn_fwd_fused_kernel<<<grid, FF_THREADS, FF_SMEM, stream>>>(tq, tk, tv, to, lse, N, Nq, H, items, scale, trace);
#ifdef TIC_ATTN_TRACE
  cudaDeviceSynchronize();
  {
    const long long t0 = trace[0];
    fprintf(stderr, "[ff trace] tile0 warp:");
    for (int i = 0; i < 32; ++i) if (trace[i]) fprintf(stderr, " a%d=%lld", i, trace[i] - t0);
    fprintf(stderr, "\n[ff trace] tile1 warp:");
    for (int i = 32; i < 64; ++i) if (trace[i]) fprintf(stderr, " b%d=%lld", i - 32, trace[i] - t0);
    fprintf(stderr, "\n[ff trace] mma thread:");
    for (int i = 64; i < 128; ++i) if (trace[i]) fprintf(stderr, " m%d=%lld", i - 64, trace[i] - t0);
    fprintf(stderr, "\n");
    cudaFree(trace);
  }
#endif
  return check_launch("attention_fwd_fused");
}

}  // namespace tic
