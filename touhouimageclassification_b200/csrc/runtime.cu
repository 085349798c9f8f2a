// Host-side runtime shared by every launcher: last-error storage, launch checking and TMA tensor-map
// encoding through the driver entry point (so the library has no link-time dependency on libcuda).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "tic_internal.cuh"

#include <cstdlib>

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

std::atomic<long long> g_launches{0};
long long launch_count() { return g_launches.load(); }

namespace {
thread_local int g_pdl_depth = 0;
}
bool pdl_enabled() {
  static const bool on = std::getenv("TIC_NO_PDL") == nullptr;
  return on && g_pdl_depth > 0;
}
PdlScope::PdlScope(bool on) : on_(on) { if (on_) ++g_pdl_depth; }
PdlScope::~PdlScope() { if (on_) --g_pdl_depth; }

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(kErrCuda, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return kOk;
}

// ---- per-device launch state. Several replicas (one model per GPU) may live in one process, each thread with its own
// current device: function attributes are per device and so is the SM count, so both are keyed by the device that is
// current at launch time (never process-global flags).
int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

int device_sm_count() {
  constexpr int kMaxDev = 64;
  static std::atomic<int> cache[kMaxDev];
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDev) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

int ensure_dynamic_smem(const void* func, int bytes, const char* what) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  const int dev = current_device();
  std::lock_guard<std::mutex> lk(mu);
  if (done.count({dev, func})) return kOk;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return set_error(kErrCuda, "%s: cudaFuncSetAttribute(%d bytes) on device %d: %s", what, bytes, dev,
                                         cudaGetErrorString(e));
  done.insert({dev, func});
  return kOk;
}

// ---- per-launch profiling (off by default): CUDA events on the launching stream around every kernel ----
namespace {
struct ProfRec { std::string name; double flops, bytes; cudaEvent_t e0, e1; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
bool g_prof_on = false;
}  // namespace

void prof_enable(bool on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_prof_on = on;
}

ProfScope::ProfScope(const char* name, double flops, double bytes, cudaStream_t s) : idx_(-1), stream_(s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{name, flops, bytes, nullptr, nullptr};
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, s);
  g_prof.push_back(r);
  idx_ = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (idx_ < static_cast<int>(g_prof.size())) cudaEventRecord(g_prof[idx_].e1, stream_);
}

// Writes "name\tlaunches\ttotal_ms\tflops\tbytes\n" per kernel name, in first-launch order. Returns bytes written.
long long prof_collect(char* buf, long long buflen) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  struct Agg { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (auto& r : g_prof) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (!agg.count(r.name)) order.push_back(r.name);
    Agg& a = agg[r.name];
    a.n += 1; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
  }
  long long off = 0;
  for (auto& name : order) {
    const Agg& a = agg[name];
    int n = snprintf(buf + off, off < buflen ? static_cast<size_t>(buflen - off) : 0, "%s\t%lld\t%.6f\t%.6e\t%.6e\n",
                     name.c_str(), a.n, a.ms, a.flops, a.bytes);
    if (n < 0 || off + n >= buflen) break;
    off += n;
  }
  return off;
}

// ---- tensor-map cache. A training step issues ~700 GEMM / attention launches whose operands sit at the same addresses
// with the same shapes every step (workspace and arenas are allocated once), so the encoded descriptors are kept in a
// thread-local table keyed by everything that goes into them; a driver call is made only on a miss. A tensor map holds
// no device state (it is 128 bytes of address arithmetic), so a cached copy stays valid as long as the key matches.
namespace {
struct TmapKey {
  uint64_t base, d0, d1, d2, p1, p2;
  uint32_t b0, b1, rank_sw;
  bool operator==(const TmapKey& o) const {
    return base == o.base && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && p1 == o.p1 && p2 == o.p2 && b0 == o.b0 && b1 == o.b1 &&
           rank_sw == o.rank_sw;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9e3779b97f4a7c15ull;
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.base); mix(k.d0); mix(k.d1); mix(k.d2); mix(k.p1); mix(k.p2); mix(k.b0); mix(k.b1); mix(k.rank_sw);
    return static_cast<size_t>(h);
  }
};
constexpr size_t kTmapCacheMax = 1 << 14;
thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> t_tmap_cache;
std::atomic<long long> g_tmap_hits{0}, g_tmap_misses{0};

bool tmap_lookup(const TmapKey& k, CUtensorMap* out) {
  auto it = t_tmap_cache.find(k);
  if (it == t_tmap_cache.end()) { g_tmap_misses.fetch_add(1, std::memory_order_relaxed); return false; }
  *out = it->second;
  g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
  return true;
}
void tmap_store(const TmapKey& k, const CUtensorMap& m) {
  if (t_tmap_cache.size() >= kTmapCacheMax) t_tmap_cache.clear();
  t_tmap_cache.emplace(k, m);
}
}  // namespace

void tmap_cache_stats(long long* hits, long long* misses) {
  *hits = g_tmap_hits.load();
  *misses = g_tmap_misses.load();
}

int encode_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t pitch_elems,
                        uint32_t box0, uint32_t box1) {
  std::call_once(g_encode_once, load_encode);
  if (g_encode == nullptr) return set_error(kErrCuda, "cuTensorMapEncodeTiled is not available from the driver");
  const TmapKey key{reinterpret_cast<uint64_t>(base), dim0, dim1, 0, pitch_elems, 0, box0, box1, 2u << 16 | 128u};
  if (tmap_lookup(key, out)) return kOk;
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
  tmap_store(key, *out);
  return kOk;
}

// 3D bf16 tensor map, 128B swizzle: [dim2][dim1][dim0] with dim0 contiguous; box = box0 x box1 x 1.
int encode_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                        uint64_t pitch1_elems, uint64_t pitch2_elems, uint32_t box0, uint32_t box1) {
  return encode_tmap_3d_bf16_sw(out, base, dim0, dim1, dim2, pitch1_elems, pitch2_elems, box0, box1, 128);
}

// Same with a selectable swizzle span in bytes (0 = none, 32, 64, 128); box0 * 2 bytes must not exceed the span.
int encode_tmap_3d_bf16_sw(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                           uint64_t pitch1_elems, uint64_t pitch2_elems, uint32_t box0, uint32_t box1, int swizzle_bytes) {
  std::call_once(g_encode_once, load_encode);
  if (g_encode == nullptr) return set_error(kErrCuda, "cuTensorMapEncodeTiled is not available from the driver");
  const TmapKey key{reinterpret_cast<uint64_t>(base), dim0, dim1, dim2, pitch1_elems, pitch2_elems, box0, box1,
                    3u << 16 | static_cast<uint32_t>(swizzle_bytes)};
  if (tmap_lookup(key, out)) return kOk;
  cuuint64_t dims[3] = {dim0, dim1, dim2};
  cuuint64_t strides[2] = {pitch1_elems * 2, pitch2_elems * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                               : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(kErrCuda, "cuTensorMapEncodeTiled(3d) failed (%d) base=%p dims=%llu,%llu,%llu box=%u,%u", (int)r, base,
                     (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)dim2, box0, box1);
  tmap_store(key, *out);
  return kOk;
}

}  // namespace tic
