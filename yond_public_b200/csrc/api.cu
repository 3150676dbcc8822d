// Library-level plumbing of libyond_b200: error reporting, launch counter, device query, tile copies.
#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include <string.h>

#include "common.cuh"

namespace {
thread_local std::string g_error;
std::atomic<uint64_t> g_launches{0};
}  // namespace

int yond_set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}
void yond_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

namespace {
struct ProfEntry {
  std::string name;
  double bytes = 0, flops = 0, ms = 0;
  long long scopes = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
std::vector<ProfEntry> g_prof;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

YondProfScope::YondProfScope(const char* name, cudaStream_t s, double bytes, double flops) : slot(-1), stream(s), end_event(nullptr) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (size_t i = 0; i < g_prof.size(); ++i)
    if (g_prof[i].name == name) slot = (int)i;
  if (slot < 0) {
    g_prof.emplace_back();
    g_prof.back().name = name;
    slot = (int)g_prof.size() - 1;
  }
  ProfEntry& e = g_prof[slot];
  e.bytes += bytes;
  e.flops += flops;
  e.scopes += 1;
  cudaEvent_t a = prof_event(), b = prof_event();
  if (!a || !b) {
    slot = -1;
    return;
  }
  cudaEventRecord(a, s);
  e.pending.emplace_back(a, b);
  end_event = b;
}
YondProfScope::~YondProfScope() {
  if (slot >= 0 && end_event) cudaEventRecord((cudaEvent_t)end_event, stream);
}

int yond_num_sms() {
  static int sms = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;
  });
  return sms;
}

namespace {
// Halo-extended tile copies on padded NHWC4 f32 frames.  Pixels outside the frame are zero: that is what the
// network's own zero padding sees at a true image border, so a tile whose halo covers the receptive field
// reproduces the whole-frame forward exactly.
__global__ void tile_extract_kernel(const float4* __restrict__ frame, float4* __restrict__ tile, int H, int W, int y0, int x0,
                                    int th, int tw) {
  const int n = th * tw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int x = i % tw + x0, y = i / tw + y0;
    tile[i] = (x >= 0 && x < W && y >= 0 && y < H) ? frame[(size_t)y * W + x] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__global__ void tile_insert_kernel(const float4* __restrict__ tile, float4* __restrict__ frame, int H, int W, int y0, int x0,
                                   int th, int tw, int halo_t, int halo_l, int core_h, int core_w) {
  const int n = core_h * core_w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int cx = i % core_w, cy = i / core_w;
    const int x = x0 + halo_l + cx, y = y0 + halo_t + cy;
    if (x >= 0 && x < W && y >= 0 && y < H) frame[(size_t)y * W + x] = tile[(size_t)(halo_t + cy) * tw + halo_l + cx];
  }
}
}  // namespace

extern "C" {

const char* yond_last_error(void) { return g_error.c_str(); }
int yond_version(void) { return 100; }
uint64_t yond_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int yond_prof_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return YOND_OK;
}

/* JSON object {"name": {"scopes": n, "ms": t, "bytes": b, "flops": f}, ...}; synchronises with the recorded events. */
int yond_prof_read(char* buf, size_t cap, int reset) {
  YOND_REQUIRE(buf && cap > 2, "yond_prof_read: no buffer");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::string out = "{";
  for (auto& e : g_prof) {
    for (auto& pr : e.pending) {
      float ms = 0.f;
      if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) e.ms += ms;
      g_prof_pool.push_back(pr.first);
      g_prof_pool.push_back(pr.second);
    }
    e.pending.clear();
    char line[512];
    snprintf(line, sizeof(line), "%s\"%s\": {\"scopes\": %lld, \"ms\": %.6f, \"bytes\": %.0f, \"flops\": %.0f}", out.size() > 1 ? ", " : "",
             e.name.c_str(), e.scopes, e.ms, e.bytes, e.flops);
    out += line;
  }
  out += "}";
  if (reset) g_prof.clear();
  YOND_REQUIRE(out.size() + 1 <= cap, "yond_prof_read: buffer too small (%zu needed)", out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return YOND_OK;
}

int yond_tile_extract(const float* frame, float* tile, int H, int W, int y0, int x0, int th, int tw, void* stream) {
  YOND_REQUIRE(th > 0 && tw > 0, "yond_tile_extract: empty tile");
  int g = ceil_div(th * tw, 256);
  if (g > yond_num_sms() * 8) g = yond_num_sms() * 8;
  tile_extract_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(frame),
                                                          reinterpret_cast<float4*>(tile), H, W, y0, x0, th, tw);
  YOND_LAUNCH_CHECK();
  return YOND_OK;
}

int yond_tile_insert(const float* tile, float* frame, int H, int W, int y0, int x0, int th, int tw, int halo_t, int halo_l,
                     int core_h, int core_w, void* stream) {
  YOND_REQUIRE(core_h > 0 && core_w > 0 && halo_t + core_h <= th && halo_l + core_w <= tw, "yond_tile_insert: core outside tile");
  int g = ceil_div(core_h * core_w, 256);
  if (g > yond_num_sms() * 8) g = yond_num_sms() * 8;
  tile_insert_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(tile), reinterpret_cast<float4*>(frame),
                                                         H, W, y0, x0, th, tw, halo_t, halo_l, core_h, core_w);
  YOND_LAUNCH_CHECK();
  return YOND_OK;
}

}  // extern "C"
