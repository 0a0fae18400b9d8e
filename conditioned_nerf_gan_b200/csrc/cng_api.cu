// Error plumbing, device check and the small host-side helpers of the C ABI (include/cng_b200.h).
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "cng_common.cuh"

namespace cng {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return CNG_OK;
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

}  // namespace cng

extern "C" {

int cng_abi_version(void) { return CNG_ABI_VERSION; }

const char* cng_last_error(void) { return cng::g_err; }

int cng_device_check(void) {
  // the verdict per device ordinal is cached after the first successful check (it is asked before every launch)
  static bool ok_cache[64] = {};
  int dev = -1;
  if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && ok_cache[dev]) return CNG_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return cng::fail(CNG_ERR_NO_DEVICE, "no CUDA device visible (%s)", e == cudaSuccess ? "count == 0" : cudaGetErrorString(e));
  }
  int major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) return cng::fail(CNG_ERR_NO_DEVICE, "device %d is sm_%d0, kernels are built for sm_100a only", dev, major);
  if (dev >= 0 && dev < 64) ok_cache[dev] = true;
  return CNG_OK;
}

// generators/volumetric_rendering.py:73-100, host part.  linspace follows torch's symmetric
// evaluation (first half from `start`, second half from `end`).
static void linspace_f32(float start, float end, int steps, float* out) {
  if (steps == 1) { out[0] = start; return; }
  const float step = (end - start) / static_cast<float>(steps - 1);
  const int half = steps / 2;
  for (int i = 0; i < steps; ++i)
    out[i] = i < half ? start + step * static_cast<float>(i) : end - step * static_cast<float>(steps - i - 1);
}

int cng_camera_tables_host(int img_w, int img_h, int S, double fov_deg, double ray_start, double ray_end,
                           float* rays_d_cam_host, float* t_lin_host) {
  CNG_REQUIRE(img_w > 0 && img_h > 0 && S > 0, CNG_ERR_INVALID_ARGUMENT, "camera_tables: non-positive size");
  CNG_REQUIRE(rays_d_cam_host && t_lin_host, CNG_ERR_INVALID_ARGUMENT, "camera_tables: NULL output");
  float* lx = new float[img_w];
  float* ly = new float[img_h];
  linspace_f32(-1.f, 1.f, img_w, lx);
  linspace_f32(-1.f, 1.f, img_h, ly);
  const float z = 1.0f / static_cast<float>(tan((2.0 * M_PI * fov_deg / 360.0) / 2.0));
  for (int r = 0; r < img_h; ++r)
    for (int c = 0; c < img_w; ++c) {
      const float x = lx[c], y = ly[r];
      const float n = sqrtf(x * x + y * y + z * z);
      float* d = rays_d_cam_host + 3 * (static_cast<size_t>(r) * img_w + c);
      d[0] = x / n; d[1] = y / n; d[2] = z / n;
    }
  linspace_f32(static_cast<float>(ray_start), static_cast<float>(ray_end), S, t_lin_host);
  delete[] lx;
  delete[] ly;
  return CNG_OK;
}

}  // extern "C"
