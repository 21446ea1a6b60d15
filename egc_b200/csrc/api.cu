// libegc_b200: error plumbing, version / build info, small utility kernels.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace egc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---- launch accounting / optional per-launch CUDA-event tracing ---------------------------------
struct ProfRecord {
  const char* name;
  cudaEvent_t start, stop;
};
static std::atomic<uint64_t> g_launches{0};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRecord> g_prof_records;

LaunchScope::LaunchScope(const char* name, cudaStream_t st) : st_(st), slot_(-1) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRecord r{name, nullptr, nullptr};
  if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
  cudaEventRecord(r.start, st);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof_records.push_back(r);
  slot_ = static_cast<int>(g_prof_records.size()) - 1;
}

LaunchScope::~LaunchScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lock(g_prof_mu);
  if (slot_ < static_cast<int>(g_prof_records.size())) cudaEventRecord(g_prof_records[slot_].stop, st_);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

__global__ void k_gather_rows(const float4* __restrict__ src, const int32_t* __restrict__ index, int n_index,
                              int width4, float4* __restrict__ dst) {
  int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  int64_t total = static_cast<int64_t>(n_index) * width4;
  if (t >= total) return;
  int r = static_cast<int>(t / width4), c = static_cast<int>(t % width4);
  dst[t] = __ldg(src + static_cast<int64_t>(index[r]) * width4 + c);
}

__global__ void k_gather_rows_scalar(const float* __restrict__ src, const int32_t* __restrict__ index, int n_index,
                                     int width, float* __restrict__ dst) {
  int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  int64_t total = static_cast<int64_t>(n_index) * width;
  if (t >= total) return;
  int r = static_cast<int>(t / width), c = static_cast<int>(t % width);
  dst[t] = __ldg(src + static_cast<int64_t>(index[r]) * width + c);
}

__global__ void k_permute_f32(const float* __restrict__ in, const int32_t* __restrict__ perm, int n,
                              float* __restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = __ldg(in + perm[t]);
}

}  // namespace egc

extern "C" {

int egc_abi_version(void) { return EGC_ABI_VERSION; }

const char* egc_last_error_string(void) { return egc::g_err; }

const char* egc_build_info(void) {
  return "libegc_b200 abi=" EGC_STR(EGC_ABI_VERSION) " arch=sm_100a cuda=" EGC_STR(__CUDACC_VER_MAJOR__) "." EGC_STR(__CUDACC_VER_MINOR__)
         " chunk_edges=" EGC_STR(EGC_CHUNK_EDGES);
}

int egc_gather_rows(const float* src, const int32_t* index, int32_t n_index, int32_t width, float* dst,
                    void* stream) {
  EGC_REQUIRE(n_index >= 0 && width > 0, "egc_gather_rows: bad sizes n_index=%d width=%d", n_index, width);
  if (n_index == 0) return EGC_OK;
  EGC_REQUIRE(src && index && dst, "egc_gather_rows: null pointer");
  const int threads = 256;
  cudaStream_t st = egc::as_stream(stream);
  bool vec = (width % 4 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) &&
             (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  egc::LaunchScope ls("k_gather_rows", st);
  if (vec) {
    int64_t total = static_cast<int64_t>(n_index) * (width / 4);
    egc::k_gather_rows<<<egc::ceil_div(total, threads), threads, 0, st>>>(
        reinterpret_cast<const float4*>(src), index, n_index, width / 4, reinterpret_cast<float4*>(dst));
  } else {
    int64_t total = static_cast<int64_t>(n_index) * width;
    egc::k_gather_rows_scalar<<<egc::ceil_div(total, threads), threads, 0, st>>>(src, index, n_index, width, dst);
  }
  EGC_LAUNCH_CHECK("k_gather_rows");
  return EGC_OK;
}

int egc_permute_f32(const float* in, const int32_t* perm, int32_t n, float* out, void* stream) {
  EGC_REQUIRE(n >= 0, "egc_permute_f32: n=%d", n);
  if (n == 0) return EGC_OK;
  EGC_REQUIRE(in && perm && out, "egc_permute_f32: null pointer");
  cudaStream_t st = egc::as_stream(stream);
  {
    egc::LaunchScope ls("k_permute_f32", st);
    egc::k_permute_f32<<<egc::ceil_div(n, 256), 256, 0, st>>>(in, perm, n, out);
  }
  EGC_LAUNCH_CHECK("k_permute_f32");
  return EGC_OK;
}

uint64_t egc_launch_count(void) { return egc::g_launches.load(); }

int egc_profile_enable(int32_t on) {
  std::lock_guard<std::mutex> lock(egc::g_prof_mu);
  for (auto& r : egc::g_prof_records) {
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  egc::g_prof_records.clear();
  egc::g_prof_on.store(on != 0);
  return EGC_OK;
}

int egc_profile_collect(char* buf, size_t capacity) {
  EGC_REQUIRE(buf != nullptr && capacity > 0, "egc_profile_collect: no buffer");
  std::lock_guard<std::mutex> lock(egc::g_prof_mu);
  std::map<std::string, std::pair<long, double>> agg;
  for (auto& r : egc::g_prof_records) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.stop) == cudaSuccess && cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
      auto& a = agg[r.name];
      a.first += 1;
      a.second += ms;
    }
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  egc::g_prof_records.clear();
  size_t off = 0;
  for (auto& kv : agg) {
    int n = snprintf(buf + off, capacity - off, "%s,%ld,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    if (n < 0 || static_cast<size_t>(n) >= capacity - off) break;
    off += static_cast<size_t>(n);
  }
  return static_cast<int>(off);
}

}  // extern "C"
