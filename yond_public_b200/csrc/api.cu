// Library-level plumbing of libyond_b200: error reporting, launch counter, device query, tile copies.
#include <atomic>
#include <mutex>
#include <string>

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
