// Fused augment + normalize + patchify: uint8 NHWC thumbnails -> bf16 patch tokens [B*196, 768] in one launch,
// replacing the reference's per-sample CPU/PIL torchvision pipeline (AugmentedDataset.setup,
// $REF/TIC/ViT/ntrain.py:104-112 [a18]) and the Conv2d im2col that follows it (modeling_vit.py:151,166 [a2]):
//   RandomResizedCrop(224) -> HFlip -> ColorJitter (random order) -> RandomGrayscale -> RandomErasing(value 0)
//   -> ToTensor -> Normalize -> 16x16 patch rows.
// Per-sample parameters (crop box, flip, jitter order + factors, gray flag, erase box) are sampled on the host from
// a counter-based hash RNG (augment_sample_params below; torchvision's sampling rules, v2/_geometry.py:272-308,
// v2/_color.py:146-171, v2/_augment.py:100-136) so a CPU restatement consumes the same integers.
// The arithmetic contract is bit-exactness against that restatement (oracle/augment_oracle.py): every float32
// operation below is an explicitly rounded intrinsic (__fmul_rn / __fadd_rn / __fdiv_rn), never a fused multiply-add.
//
// A cluster of two CTAs per image: each owns half of the output rows (whole 16-row patch bands) and keeps its part of the
// uint8 working image in shared memory (planar, <= 74 KB), so the colour ops never round-trip through HBM and two CTAs
// (32 warps) share an SM. The one whole-image quantity, the grey mean of the contrast step, is summed across the pair
// through distributed shared memory. ToTensor + Normalize of a uint8 pixel takes 256 values per channel: they are
// tabulated once per CTA with the reference's own rounded operations, so the output passes are table look-ups.
// Algorithmic HBM traffic per image: <= H*W*3 bytes read (the crop), 196*768*2 = 301056 bytes written.
#include <cmath>

#include "tic_internal.cuh"

namespace tic {
namespace {

constexpr int AUG_THREADS = 512;
constexpr int AUG_MAX_TAPS = 8;
constexpr int AUG_MAX_SIZE = 224;

struct AugTables {
  int lo[2][AUG_MAX_SIZE];
  int n[2][AUG_MAX_SIZE];
  float w[2][AUG_MAX_SIZE][AUG_MAX_TAPS];
};

__device__ __forceinline__ float gray_floor_f(float r, float g, float b) {
  // floor(0.2989 r + 0.587 g + 0.114 b), products and sums separately rounded
  return floorf(__fadd_rn(__fadd_rn(__fmul_rn(r, 0.2989f), __fmul_rn(g, 0.587f)), __fmul_rn(b, 0.114f)));
}
__device__ __forceinline__ unsigned char to_u8(float x) {  // clamp to [0, 255] then truncate
  return static_cast<unsigned char>(fminf(fmaxf(x, 0.0f), 255.0f));
}
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

constexpr int AUG_CLUSTER = 2;                                   // CTAs per image
constexpr int AUG_MAX_ROWS = ((AUG_MAX_SIZE / 16 + 1) / 2) * 16;  // output rows per CTA (112)

// this CTA's and the peer's value of a shared-memory word (distributed shared memory)
__device__ __forceinline__ int dsmem_read_peer(const int* p, unsigned peer) {
  unsigned local = static_cast<unsigned>(__cvta_generic_to_shared(p)), remote;
  int v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(peer));
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
  return v;
}

__global__ void __launch_bounds__(AUG_THREADS, 2)
augment_patchify_kernel(const unsigned char* __restrict__ images, int H, int W, const int* __restrict__ ints,
                        const float* __restrict__ floats, int size, float m0, float m1, float m2, float s0, float s1,
                        float s2, __nv_bfloat16* __restrict__ patches, unsigned char* __restrict__ pixels_out,
                        float* __restrict__ tensor_out) {
  extern __shared__ __align__(16) unsigned char aug_smem[];
  AugTables* tab = reinterpret_cast<AugTables*>(aug_smem);
  float* lut = reinterpret_cast<float*>(aug_smem + ((sizeof(AugTables) + 15) / 16) * 16);  // [3][256] normalised values
  unsigned char* pix = reinterpret_cast<unsigned char*>(lut + 3 * 256);                    // planar [3][rows * size]
  __shared__ int gray_sum;
  const unsigned rank = cluster_ctarank();
  const int b = blockIdx.x / AUG_CLUSTER, tid = threadIdx.x;
  const int* pi = ints + b * 16;
  const float* pf = floats + b * 4;
  const int top = pi[0], left = pi[1], ch = pi[2], cw = pi[3], flip = pi[4];
  const int G = size / 16;
  // this CTA's patch bands [g0, g0 + ng) = output rows [y0, y0 + rows)
  const int g0 = rank == 0 ? 0 : (G + 1) / 2, ng = rank == 0 ? (G + 1) / 2 : G - (G + 1) / 2;
  const int y0 = g0 * 16, rows = ng * 16;
  const int npix = rows * size;          // pixels of this CTA
  const int plane = npix;                // plane pitch of the working image
  const int npix_image = size * size;
  const unsigned char* src = images + static_cast<long long>(b) * H * W * 3;
  if (tid == 0) gray_sum = 0;

  // ---- 0. ToTensor + Normalize of every uint8 value, per channel: (v / 255 - mean) / std, each operation rounded
  {
    const float mean_c[3] = {m0, m1, m2}, std_c[3] = {s0, s1, s2};
    for (int e = tid; e < 3 * 256; e += AUG_THREADS) {
      const int c = e >> 8;
      lut[e] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(e & 255), 255.0f), mean_c[c]), std_c[c]);
    }
  }

  // ---- 1. antialiased-bilinear tap tables (x: dir 0 over crop width, y: dir 1 over crop height)
  for (int e = tid; e < 2 * size; e += AUG_THREADS) {
    const int dir = e / size, o = e - dir * size;
    const int in = dir == 0 ? cw : ch;
    const float scale = __fdiv_rn(static_cast<float>(in), static_cast<float>(size));
    const float fs = fmaxf(scale, 1.0f);
    const float center = __fmul_rn(__fadd_rn(static_cast<float>(o), 0.5f), scale);
    int a = static_cast<int>(__fadd_rn(__fsub_rn(center, fs), 0.5f));
    int e2 = static_cast<int>(__fadd_rn(__fadd_rn(center, fs), 0.5f));
    a = max(a, 0);
    e2 = min(e2, in);
    const int cnt = min(e2 - a, AUG_MAX_TAPS);
    float raw[AUG_MAX_TAPS];
    float total = 0.0f;
#pragma unroll
    for (int k = 0; k < AUG_MAX_TAPS; ++k) {
      float wk = 0.0f;
      if (k < cnt) {
        const float x = __fdiv_rn(__fadd_rn(__fsub_rn(static_cast<float>(a + k), center), 0.5f), fs);
        wk = fmaxf(0.0f, __fsub_rn(1.0f, fabsf(x)));
        total = __fadd_rn(total, wk);
      }
      raw[k] = wk;
    }
#pragma unroll
    for (int k = 0; k < AUG_MAX_TAPS; ++k) tab->w[dir][o][k] = k < cnt ? __fdiv_rn(raw[k], total) : 0.0f;
    tab->lo[dir][o] = a;
    tab->n[dir][o] = cnt;
  }
  __syncthreads();

  // ---- 2. resized crop (+ horizontal flip) -> uint8 working image in shared memory (this CTA's rows)
  for (int p = tid; p < npix; p += AUG_THREADS) {
    const int yl = p / size, xo = p - yl * size;
    const int yo = y0 + yl;
    const int xr = flip ? size - 1 - xo : xo;
    const int lox = tab->lo[0][xr], nx = tab->n[0][xr], loy = tab->lo[1][yo], ny = tab->n[1][yo];
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    for (int ty = 0; ty < ny; ++ty) {
      const unsigned char* row = src + (static_cast<long long>(top + loy + ty) * W + left + lox) * 3;
      float r0 = 0.f, r1 = 0.f, r2 = 0.f;
      for (int tx = 0; tx < nx; ++tx) {
        const float wx = tab->w[0][xr][tx];
        r0 = __fadd_rn(r0, __fmul_rn(static_cast<float>(__ldg(row + 3 * tx + 0)), wx));
        r1 = __fadd_rn(r1, __fmul_rn(static_cast<float>(__ldg(row + 3 * tx + 1)), wx));
        r2 = __fadd_rn(r2, __fmul_rn(static_cast<float>(__ldg(row + 3 * tx + 2)), wx));
      }
      const float wy = tab->w[1][yo][ty];
      acc0 = __fadd_rn(acc0, __fmul_rn(r0, wy));
      acc1 = __fadd_rn(acc1, __fmul_rn(r1, wy));
      acc2 = __fadd_rn(acc2, __fmul_rn(r2, wy));
    }
    pix[p] = to_u8(floorf(__fadd_rn(acc0, 0.5f)));
    pix[plane + p] = to_u8(floorf(__fadd_rn(acc1, 0.5f)));
    pix[2 * plane + p] = to_u8(floorf(__fadd_rn(acc2, 0.5f)));
  }
  // each thread keeps working on its own pixels: no barrier needed until the contrast mean

  // ---- 3. ColorJitter in the sampled order (torchvision tensor kernels on uint8 images)
  if (pi[14]) {
    for (int k = 0; k < 4; ++k) {
      const int op = pi[5 + k];
      if (op == 0) {  // brightness: clamp(x * b)
        const float f = pf[0];
        for (int p = tid; p < npix; p += AUG_THREADS) {
#pragma unroll
          for (int c = 0; c < 3; ++c) pix[c * plane + p] = to_u8(__fmul_rn(static_cast<float>(pix[c * plane + p]), f));
        }
      } else if (op == 1) {  // contrast: blend(x, mean(gray), c); the mean is over the WHOLE image: both CTAs of the pair
        int local = 0;
        for (int p = tid; p < npix; p += AUG_THREADS)
          local += static_cast<int>(gray_floor_f(pix[p], pix[plane + p], pix[2 * plane + p]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((tid & 31) == 0) atomicAdd(&gray_sum, local);
        cluster_sync_all();   // (also a CTA barrier) both partial sums are complete and visible across the pair
        const int total = gray_sum + dsmem_read_peer(&gray_sum, rank ^ 1u);
        const float mean = __fdiv_rn(static_cast<float>(total), static_cast<float>(npix_image));
        const float f = pf[1], omf = __fsub_rn(1.0f, f);
        const float mterm = __fmul_rn(mean, omf);
        for (int p = tid; p < npix; p += AUG_THREADS) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            pix[c * plane + p] = to_u8(__fadd_rn(__fmul_rn(static_cast<float>(pix[c * plane + p]), f), mterm));
        }
      } else if (op == 2) {  // saturation: blend(x, gray(x), s)
        const float f = pf[2], omf = __fsub_rn(1.0f, f);
        for (int p = tid; p < npix; p += AUG_THREADS) {
          const float r = pix[p], g = pix[plane + p], bl = pix[2 * plane + p];
          const float gterm = __fmul_rn(gray_floor_f(r, g, bl), omf);
          pix[p] = to_u8(__fadd_rn(__fmul_rn(r, f), gterm));
          pix[plane + p] = to_u8(__fadd_rn(__fmul_rn(g, f), gterm));
          pix[2 * plane + p] = to_u8(__fadd_rn(__fmul_rn(bl, f), gterm));
        }
      } else {  // hue: RGB -> HSV, shift h, HSV -> RGB (v2/functional/_color.py:300-400)
        const float hf = pf[3];
        for (int p = tid; p < npix; p += AUG_THREADS) {
          const float inv255 = 0.00392156862745098f;
          const float r = __fmul_rn(static_cast<float>(pix[p]), inv255);
          const float g = __fmul_rn(static_cast<float>(pix[plane + p]), inv255);
          const float bl = __fmul_rn(static_cast<float>(pix[2 * plane + p]), inv255);
          const float maxc = fmaxf(fmaxf(r, g), bl), minc = fminf(fminf(r, g), bl);
          const bool eqc = maxc == minc;
          const float cr = __fsub_rn(maxc, minc);
          const float s = __fdiv_rn(cr, eqc ? 1.0f : maxc);
          const float dv = eqc ? 1.0f : cr;
          const float rc = __fdiv_rn(__fsub_rn(maxc, r), dv), gc = __fdiv_rn(__fsub_rn(maxc, g), dv),
                      bc = __fdiv_rn(__fsub_rn(maxc, bl), dv);
          const bool neq_r = maxc != r, eq_g = maxc == g;
          const float hg = __fmul_rn(__fsub_rn(__fadd_rn(rc, 2.0f), bc), (eq_g && neq_r) ? 1.0f : 0.0f);
          const float hr = __fmul_rn(__fsub_rn(bc, gc), neq_r ? 0.0f : 1.0f);
          const float hb = __fmul_rn(__fsub_rn(__fadd_rn(gc, 4.0f), rc), (neq_r && !eq_g) ? 1.0f : 0.0f);
          float h = __fadd_rn(__fadd_rn(hr, hg), hb);
          h = fmodf(__fadd_rn(__fmul_rn(h, 0.16666666666666666f), 1.0f), 1.0f);
          h = fmodf(__fadd_rn(h, hf), 1.0f);
          if (h < 0.0f) h = __fadd_rn(h, 1.0f);
          const float v = maxc;
          const float h6 = __fmul_rn(h, 6.0f);
          const float fi = floorf(h6);
          const float f = __fsub_rn(h6, fi);
          const int i = static_cast<int>(fi) % 6;
          const float sxf = __fmul_rn(s, f);
          const float oms = __fsub_rn(1.0f, s);
          const float q = clamp01(__fmul_rn(__fsub_rn(1.0f, sxf), v));
          const float t = clamp01(__fmul_rn(__fadd_rn(sxf, oms), v));
          const float pp = clamp01(__fmul_rn(oms, v));
          float ro, go, bo;
          switch (i) {
            case 0: ro = v; go = t; bo = pp; break;
            case 1: ro = q; go = v; bo = pp; break;
            case 2: ro = pp; go = v; bo = t; break;
            case 3: ro = pp; go = q; bo = v; break;
            case 4: ro = t; go = pp; bo = v; break;
            default: ro = v; go = pp; bo = q; break;
          }
          pix[p] = static_cast<unsigned char>(__fmul_rn(ro, 255.999f));
          pix[plane + p] = static_cast<unsigned char>(__fmul_rn(go, 255.999f));
          pix[2 * plane + p] = static_cast<unsigned char>(__fmul_rn(bo, 255.999f));
        }
      }
    }
  }

  // ---- 4. RandomGrayscale, RandomErasing (value 0, before normalisation)
  const int gray = pi[9], ei = pi[10], ej = pi[11], eh = pi[12], ew = pi[13];
  if (gray || (eh > 0 && ew > 0) || pixels_out != nullptr) {
    for (int p = tid; p < npix; p += AUG_THREADS) {
      unsigned char r = pix[p], g = pix[plane + p], bl = pix[2 * plane + p];
      if (gray) r = g = bl = static_cast<unsigned char>(gray_floor_f(r, g, bl));
      const int yl = p / size, x = p - yl * size;
      const int y = y0 + yl;
      if (y >= ei && y < ei + eh && x >= ej && x < ej + ew) r = g = bl = 0;
      pix[p] = r; pix[plane + p] = g; pix[2 * plane + p] = bl;
      if (pixels_out != nullptr) {
        unsigned char* o = pixels_out + (static_cast<long long>(b) * npix_image + y0 * size + p) * 3;
        o[0] = r; o[1] = g; o[2] = bl;
      }
    }
  }
  __syncthreads();

  // ---- 5. ToTensor (/255), Normalize (table look-ups), bf16, patch rows with K ordered (c, py, px); 16-byte stores
  if (tensor_out != nullptr) {  // the fp32 [3, size, size] tensor the reference's Dataset yields (what CutMix / MixUp blend)
    float* to = tensor_out + static_cast<long long>(b) * 3 * npix_image + y0 * size;
    for (int e = tid; e < 3 * npix / 4; e += AUG_THREADS) {
      const int c = (e * 4) / npix;  // npix is a multiple of 256: a group of four never straddles a plane
      const int q = e * 4 - c * npix;
      const unsigned int raw = *reinterpret_cast<const unsigned int*>(pix + c * plane + q);
      const float* l = lut + c * 256;
      *reinterpret_cast<float4*>(to + static_cast<long long>(c) * npix_image + q) =
          make_float4(l[raw & 0xffu], l[(raw >> 8) & 0xffu], l[(raw >> 16) & 0xffu], l[raw >> 24]);
    }
  }
  if (patches != nullptr) {
    __nv_bfloat16* out = patches + (static_cast<long long>(b) * G + g0) * G * 768;
    for (int e = tid; e < ng * G * 96; e += AUG_THREADS) {
      const int row = e / 96, chunk = e - row * 96;   // row = local patch index (band-major)
      const int k = chunk * 8;
      const int c = k >> 8, py = (k & 255) >> 4, px = k & 15;
      const int gy = row / G, gx = row - gy * G;
      const unsigned char* sp = pix + c * plane + (gy * 16 + py) * size + gx * 16 + px;
      const uint2 raw = *reinterpret_cast<const uint2*>(sp);
      const float* l = lut + c * 256;
      uint4 w;
      w.x = pack_bf16x2(l[raw.x & 0xffu], l[(raw.x >> 8) & 0xffu]);
      w.y = pack_bf16x2(l[(raw.x >> 16) & 0xffu], l[raw.x >> 24]);
      w.z = pack_bf16x2(l[raw.y & 0xffu], l[(raw.y >> 8) & 0xffu]);
      w.w = pack_bf16x2(l[(raw.y >> 16) & 0xffu], l[raw.y >> 24]);
      *reinterpret_cast<uint4*>(out + static_cast<long long>(row) * 768 + k) = w;
    }
  }
  cluster_sync_all();  // neither CTA of the pair exits while the other may still read its grey sum
}

// ---------------------------------------------------------------------------------------------- host sampler
inline uint32_t rng_u32(uint64_t seed, uint64_t sample, uint64_t draw) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + sample * 0xBF58476D1CE4E5B9ull + draw * 0x94D049BB133111EBull +
               0x2545F4914F6CDD1Dull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<uint32_t>(z >> 32);
}
struct Stream {
  uint64_t seed, sample, draw;
  double uniform(double lo, double hi) {
    const double u = static_cast<double>(rng_u32(seed, sample, draw++) >> 8) * (1.0 / 16777216.0);
    return lo + (hi - lo) * u;
  }
  int randint(int n) { return static_cast<int>(rng_u32(seed, sample, draw++) % static_cast<uint32_t>(n)); }
};
inline int round_half_even(double x) { return static_cast<int>(std::nearbyint(x)); }  // == Python round()

}  // namespace

// recipe: 0 = full (ntrain.py:104-112), 1 = generalization only (crop + flip + erase, :127-134), 2 = diversity only
// (Resize + ColorJitter + RandomGrayscale, :119-126), 3 = grey only (Resize + RandomGrayscale, :97-102),
// 4 = none (Resize only: the test / inference transform, :137-147 and utils/preprocess.py:73-77).
// Random numbers are drawn only for the transforms a recipe contains, in pipeline order.
int augment_sample_params(long long seed, long long first_sample, int B, int H, int W, int size, int recipe,
                          int* ints_host, float* floats_host) {
  if (B < 0 || H <= 0 || W <= 0 || size <= 0) return set_error(kErrInvalidArg, "augment_sample_params: bad sizes");
  if (recipe < 0 || recipe > 4) return set_error(kErrInvalidArg, "augment_sample_params: unknown recipe %d", recipe);
  const bool do_crop = recipe <= 1, do_flip = recipe <= 1, do_jitter = recipe == 0 || recipe == 2;
  const bool do_gray = recipe == 0 || recipe == 2 || recipe == 3, do_erase = recipe <= 1;
  for (int b = 0; b < B; ++b) {
    Stream st{static_cast<uint64_t>(seed), static_cast<uint64_t>(first_sample + b), 0};
    int* I = ints_host + 16 * b;
    float* F = floats_host + 4 * b;
    const double area = static_cast<double>(H) * W;
    const double lr0 = std::log(3.0 / 4.0), lr1 = std::log(4.0 / 3.0);
    int top = 0, left = 0, h = H, w = W;
    bool found = !do_crop;  // Resize-only recipes use the whole frame
    for (int t = 0; t < 10 && !found; ++t) {
      const double target = area * st.uniform(0.08, 1.0);
      const double aspect = std::exp(st.uniform(lr0, lr1));
      const int cw = round_half_even(std::sqrt(target * aspect));
      const int chh = round_half_even(std::sqrt(target / aspect));
      if (0 < cw && cw <= W && 0 < chh && chh <= H) {
        top = st.randint(H - chh + 1);
        left = st.randint(W - cw + 1);
        h = chh; w = cw;
        found = true;
      }
    }
    if (!found && do_crop) {
      const double in_ratio = static_cast<double>(W) / static_cast<double>(H);
      if (in_ratio < 3.0 / 4.0) { w = W; h = round_half_even(w / (3.0 / 4.0)); }
      else if (in_ratio > 4.0 / 3.0) { h = H; w = round_half_even(h * (4.0 / 3.0)); }
      else { w = W; h = H; }
      top = (H - h) / 2; left = (W - w) / 2;
    }
    const int flip = do_flip ? (st.uniform(0.0, 1.0) < 0.5 ? 1 : 0) : 0;
    int perm[4] = {0, 1, 2, 3};
    for (int i = 3; i > 0 && do_jitter; --i) {
      const int j = st.randint(i + 1);
      const int tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
    }
    double bb = 1.0, cc = 1.0, ss = 1.0, hh = 0.0;
    if (do_jitter) { bb = st.uniform(0.8, 1.2); cc = st.uniform(0.8, 1.2); ss = st.uniform(0.8, 1.2); hh = st.uniform(-0.1, 0.1); }
    const int gray = do_gray ? (st.uniform(0.0, 1.0) < 0.2 ? 1 : 0) : 0;
    int ei = 0, ej = 0, eh = 0, ew = 0;
    if (do_erase && st.uniform(0.0, 1.0) < 0.5) {
      const double el0 = std::log(0.3), el1 = std::log(3.3);
      for (int t = 0; t < 10; ++t) {
        const double ea = static_cast<double>(size) * size * st.uniform(0.02, 0.33);
        const double aspect = std::exp(st.uniform(el0, el1));
        const int rh = round_half_even(std::sqrt(ea * aspect));
        const int rw = round_half_even(std::sqrt(ea / aspect));
        if (!(rh < size && rw < size)) continue;
        ei = st.randint(size - rh + 1);
        ej = st.randint(size - rw + 1);
        eh = rh; ew = rw;
        break;
      }
    }
    const int jitter_on = do_jitter ? 1 : 0;
    const int vals[16] = {top, left, h, w, flip, perm[0], perm[1], perm[2], perm[3], gray, ei, ej, eh, ew, jitter_on, 0};
    for (int i = 0; i < 16; ++i) I[i] = vals[i];
    F[0] = static_cast<float>(bb); F[1] = static_cast<float>(cc); F[2] = static_cast<float>(ss); F[3] = static_cast<float>(hh);
  }
  return kOk;
}

int augment_patchify(const void* images_u8, int B, int H, int W, const int* ints_dev, const float* floats_dev, int size,
                     const float* mean3_host, const float* std3_host, void* patches_bf16, void* pixels_out_u8,
                     float* tensor_out_f32, cudaStream_t stream) {
  if (size % 16 != 0 || size > AUG_MAX_SIZE || size <= 0)
    return set_error(kErrUnsupported, "augment_patchify: output size %d (multiple of 16, <= %d)", size, AUG_MAX_SIZE);
  if (B <= 0) return kOk;
  if (patches_bf16 == nullptr && tensor_out_f32 == nullptr)
    return set_error(kErrInvalidArg, "augment_patchify: no output requested (patches and tensor are both NULL)");
  // tap tables hold at most AUG_MAX_TAPS taps: support = max(in/out, 1) must be <= 3.5
  if (H > 3 * size || W > 3 * size)
    return set_error(kErrUnsupported, "augment_patchify: source %dx%d is more than 3x the output size", H, W);
  const int lut_bytes = 3 * 256 * 4;
  const int rows = ((size / 16 + 1) / 2) * 16;   // output rows of the larger half
  const int smem = static_cast<int>((sizeof(AugTables) + 15) / 16 * 16) + lut_bytes + 3 * rows * size;
  if (int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(augment_patchify_kernel),
                                   static_cast<int>((sizeof(AugTables) + 15) / 16 * 16) + lut_bytes +
                                       3 * AUG_MAX_ROWS * AUG_MAX_SIZE,
                                   "augment_patchify"))
    return rc;
  ProfScope prof("augment_patchify", 0.0, static_cast<double>(B) * (static_cast<double>(H) * W * 3 + (size / 16) * (size / 16) * 768.0 * 2), stream);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(B * AUG_CLUSTER);
  cfg.blockDim = dim3(AUG_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = AUG_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, augment_patchify_kernel, reinterpret_cast<const unsigned char*>(images_u8), H, W,
                                     ints_dev, floats_dev, size, mean3_host[0], mean3_host[1], mean3_host[2], std3_host[0],
                                     std3_host[1], std3_host[2], reinterpret_cast<__nv_bfloat16*>(patches_bf16),
                                     reinterpret_cast<unsigned char*>(pixels_out_u8), tensor_out_f32);
  if (e != cudaSuccess) return set_error(kErrCuda, "augment_patchify: cudaLaunchKernelEx: %s", cudaGetErrorString(e));
  return check_launch("augment_patchify");
}

}  // namespace tic
