// GroupNorm over channels-last volumes [N, S, C] (S = D*H*W), forward and backward: the normalisation of the 3D U-Net encoder
// that produces the feature volume the rendering path reads (generators/unet3d.py:21-132, nn.GroupNorm
// inside every SingleConv, order "gcr").
//
// Why a kernel of our own: the encoder runs in channels_last_3d so that its last convolution emits the NDHWC volume K1 gathers
// from, and ATen's native_group_norm reads such a tensor with a stride of S elements per channel (RowwiseMomentsCUDAKernel:
// 0.44 ms per call, 13 % of a batch-4 train step) after autocast has widened it to fp32.  Here a group's channels are contiguous
// in memory, every access is a 16-byte vector, 16-bit tensors are read and written as they are (statistics in fp32 / fp64):
//   forward   pass 1: per (n, g) sum and sum of squares (fp32 per thread and block, fp64 atomics across blocks)
//             pass 2: y = x * a[n,c] + b[n,c] with a = rstd * gamma, b = beta - mean * a; mean / rstd [N, G] kept for backward
//   backward  pass 1: per (n, c) sum dy * xhat and sum dy  (-> d_gamma, d_beta, and the two per-group sums of the input gradient)
//             pass 2: dx = rstd * (dy * gamma - (xhat * c1 + c2) / m),  c1 = sum_g dy gamma xhat, c2 = sum_g dy gamma, m = S * C/G
// Algorithmic bytes: forward 2 reads + 1 write of the tensor, backward 2 reads of (dy, x) + 1 write.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>

#include "cng_common.cuh"

namespace cng {
namespace gn {

constexpr int kThreads = 256;
constexpr int kMaxC = 1024;

template <typename T> struct Vec;           // 16-byte vectors
template <> struct Vec<float> { static constexpr int n = 4; };
template <> struct Vec<__half> { static constexpr int n = 8; };
template <> struct Vec<__nv_bfloat16> { static constexpr int n = 8; };

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// V elements per access (V divides C/G, so a vector never straddles a group; V * sizeof(T) <= 16)
template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* p, float (&out)[V]) {
  if constexpr (V * sizeof(T) == 16) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const T* e = reinterpret_cast<const T*>(&r);
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = to_f<T>(e[i]);
  } else if constexpr (V * sizeof(T) == 8) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const T* e = reinterpret_cast<const T*>(&r);
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = to_f<T>(e[i]);
  } else if constexpr (V * sizeof(T) == 4) {
    const uint32_t r = __ldg(reinterpret_cast<const uint32_t*>(p));
    const T* e = reinterpret_cast<const T*>(&r);
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = to_f<T>(e[i]);
  } else {
#pragma unroll
    for (int i = 0; i < V; ++i) out[i] = to_f<T>(p[i]);
  }
}
template <typename T, int V>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[V]) {
  T e[V];
#pragma unroll
  for (int i = 0; i < V; ++i) e[i] = from_f<T>(v[i]);
  if constexpr (V * sizeof(T) == 16) *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(e);
  else if constexpr (V * sizeof(T) == 8) *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(e);
  else if constexpr (V * sizeof(T) == 4) *reinterpret_cast<uint32_t*>(p) = *reinterpret_cast<const uint32_t*>(e);
  else {
#pragma unroll
    for (int i = 0; i < V; ++i) p[i] = e[i];
  }
}

struct Params {
  const void* x;
  const void* dy;
  void* y;              // forward: y; backward: dx
  long long S;
  int N, C, G;
  const float* gamma;   // [C] or NULL (1)
  const float* beta;    // [C] or NULL (0)
  float eps;
  float* mean;          // [N, G]
  float* rstd;          // [N, G]
  double* sums;         // forward scratch [N, G, 2], zeroed by the caller
  float* ds;            // backward scratch [N, C]: sum dy * xhat, zeroed by the caller
  float* db;            // backward scratch [N, C]: sum dy
  long long rows_per_block;
};

// ---- forward pass 1: sums per (n, g) -------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads) stats_kernel(Params p) {
  __shared__ float s_sum[64], s_sq[64];
  const int n = blockIdx.y, C = p.C, CV = C / V, cpg = C / p.G;
  if (threadIdx.x < 64) { s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; }
  __syncthreads();
  const int rows_par = kThreads / CV;                       // positions handled per sweep of the block
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  const long long s0 = blockIdx.x * p.rows_per_block, s1 = min(p.S, s0 + p.rows_per_block);
  float a = 0.f, q = 0.f;
  if (r < rows_par) {
    const T* x = static_cast<const T*>(p.x) + (static_cast<size_t>(n) * p.S) * C + cv * V;
    for (long long s = s0 + r; s < s1; s += rows_par) {
      float v[V];
      load_vec<T, V>(x + s * C, v);
#pragma unroll
      for (int i = 0; i < V; ++i) { a += v[i]; q = fmaf(v[i], v[i], q); }
    }
    const int g = (cv * V) / cpg;
    atomicAdd(&s_sum[g], a);
    atomicAdd(&s_sq[g], q);
  }
  __syncthreads();
  if (threadIdx.x < p.G) {
    atomicAdd(p.sums + (static_cast<size_t>(n) * p.G + threadIdx.x) * 2, static_cast<double>(s_sum[threadIdx.x]));
    atomicAdd(p.sums + (static_cast<size_t>(n) * p.G + threadIdx.x) * 2 + 1, static_cast<double>(s_sq[threadIdx.x]));
  }
}

// ---- forward pass 2: y = x * a + b ----------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads) apply_kernel(Params p) {
  __shared__ float s_a[kMaxC], s_b[kMaxC];
  const int n = blockIdx.y, C = p.C, CV = C / V, cpg = C / p.G;
  const double m = static_cast<double>(p.S) * cpg;
  for (int c = threadIdx.x; c < C; c += kThreads) {
    const int g = c / cpg;
    const double su = p.sums[(static_cast<size_t>(n) * p.G + g) * 2], sq = p.sums[(static_cast<size_t>(n) * p.G + g) * 2 + 1];
    const double mean = su / m;
    const double var = fmax(sq / m - mean * mean, 0.0);
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
    const float ga = p.gamma ? __ldg(p.gamma + c) : 1.f, be = p.beta ? __ldg(p.beta + c) : 0.f;
    s_a[c] = rstd * ga;
    s_b[c] = be - static_cast<float>(mean) * rstd * ga;
    if (blockIdx.x == 0 && c % cpg == 0) {
      p.mean[static_cast<size_t>(n) * p.G + g] = static_cast<float>(mean);
      p.rstd[static_cast<size_t>(n) * p.G + g] = rstd;
    }
  }
  __syncthreads();
  const int rows_par = kThreads / CV;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  if (r >= rows_par) return;
  const long long s0 = blockIdx.x * p.rows_per_block, s1 = min(p.S, s0 + p.rows_per_block);
  const T* x = static_cast<const T*>(p.x) + (static_cast<size_t>(n) * p.S) * C + cv * V;
  T* y = static_cast<T*>(p.y) + (static_cast<size_t>(n) * p.S) * C + cv * V;
  float a[V], b[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { a[i] = s_a[cv * V + i]; b[i] = s_b[cv * V + i]; }
  for (long long s = s0 + r; s < s1; s += rows_par) {
    float v[V];
    load_vec<T, V>(x + s * C, v);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = fmaf(v[i], a[i], b[i]);
    store_vec<T, V>(y + s * C, v);
  }
}

// ---- backward pass 1: ds[n,c] = sum_s dy * xhat, db[n,c] = sum_s dy ---------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads) bwd_stats_kernel(Params p) {
  __shared__ float s_ds[kMaxC], s_db[kMaxC];
  const int n = blockIdx.y, C = p.C, CV = C / V, cpg = C / p.G;
  for (int c = threadIdx.x; c < C; c += kThreads) { s_ds[c] = 0.f; s_db[c] = 0.f; }
  __syncthreads();
  const int rows_par = kThreads / CV;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  if (r < rows_par) {
    const int g = (cv * V) / cpg;
    const float mean = __ldg(p.mean + static_cast<size_t>(n) * p.G + g), rstd = __ldg(p.rstd + static_cast<size_t>(n) * p.G + g);
    const long long s0 = blockIdx.x * p.rows_per_block, s1 = min(p.S, s0 + p.rows_per_block);
    const T* x = static_cast<const T*>(p.x) + (static_cast<size_t>(n) * p.S) * C + cv * V;
    const T* dy = static_cast<const T*>(p.dy) + (static_cast<size_t>(n) * p.S) * C + cv * V;
    float a[V], b[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { a[i] = 0.f; b[i] = 0.f; }
    for (long long s = s0 + r; s < s1; s += rows_par) {
      float v[V], d[V];
      load_vec<T, V>(x + s * C, v);
      load_vec<T, V>(dy + s * C, d);
#pragma unroll
      for (int i = 0; i < V; ++i) { a[i] = fmaf(d[i], (v[i] - mean) * rstd, a[i]); b[i] += d[i]; }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) { atomicAdd(&s_ds[cv * V + i], a[i]); atomicAdd(&s_db[cv * V + i], b[i]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kThreads) {
    atomicAdd(p.ds + static_cast<size_t>(n) * C + c, s_ds[c]);
    atomicAdd(p.db + static_cast<size_t>(n) * C + c, s_db[c]);
  }
}

// ---- backward pass 2: dx ---------------------------------------------------------------------------------------------------
template <typename T, int V>
__global__ void __launch_bounds__(kThreads) bwd_apply_kernel(Params p) {
  __shared__ float s_a[kMaxC];           // rstd * gamma per channel
  __shared__ float s_k1[64], s_k2[64];   // per group: dx = dy * a[c] + x * k1[g] + k2[g]
  const int n = blockIdx.y, C = p.C, CV = C / V, cpg = C / p.G;
  const float inv_m = 1.f / (static_cast<float>(p.S) * cpg);
  if (threadIdx.x < p.G) {
    const int g = threadIdx.x;
    const float mean = __ldg(p.mean + static_cast<size_t>(n) * p.G + g), rstd = __ldg(p.rstd + static_cast<size_t>(n) * p.G + g);
    float c1 = 0.f, c2 = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const float ga = p.gamma ? __ldg(p.gamma + c) : 1.f;
      c1 = fmaf(ga, __ldg(p.ds + static_cast<size_t>(n) * C + c), c1);
      c2 = fmaf(ga, __ldg(p.db + static_cast<size_t>(n) * C + c), c2);
    }
    // dx = rstd * (dy * gamma - (xhat * c1 + c2) / m),  xhat = (x - mean) * rstd
    const float k1 = -rstd * rstd * c1 * inv_m;
    s_k1[g] = k1;
    s_k2[g] = -mean * k1 - rstd * c2 * inv_m;
  }
  for (int c = threadIdx.x; c < C; c += kThreads)
    s_a[c] = __ldg(p.rstd + static_cast<size_t>(n) * p.G + c / cpg) * (p.gamma ? __ldg(p.gamma + c) : 1.f);
  __syncthreads();
  const int rows_par = kThreads / CV;
  const int cv = threadIdx.x % CV, r = threadIdx.x / CV;
  if (r >= rows_par) return;
  const int g = (cv * V) / cpg;
  const float k1 = s_k1[g], k2 = s_k2[g];
  float a[V];
#pragma unroll
  for (int i = 0; i < V; ++i) a[i] = s_a[cv * V + i];
  const long long s0 = blockIdx.x * p.rows_per_block, s1 = min(p.S, s0 + p.rows_per_block);
  const T* x = static_cast<const T*>(p.x) + (static_cast<size_t>(n) * p.S) * C + cv * V;
  const T* dy = static_cast<const T*>(p.dy) + (static_cast<size_t>(n) * p.S) * C + cv * V;
  T* dx = static_cast<T*>(p.y) + (static_cast<size_t>(n) * p.S) * C + cv * V;
  for (long long s = s0 + r; s < s1; s += rows_par) {
    float v[V], d[V];
    load_vec<T, V>(x + s * C, v);
    load_vec<T, V>(dy + s * C, d);
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = fmaf(d[i], a[i], fmaf(v[i], k1, k2));
    store_vec<T, V>(dx + s * C, v);
  }
}

template <typename T, int V>
static int launch(Params& p, bool backward, cudaStream_t st) {
  // enough blocks for a few waves, at least 64 positions per block
  const long long want = std::max<long long>(1, (8LL * sm_count()) / std::max(1, p.N));
  p.rows_per_block = std::max<long long>(64, (p.S + want - 1) / want);
  dim3 grid(static_cast<unsigned>((p.S + p.rows_per_block - 1) / p.rows_per_block), p.N);
  if (!backward) {
    stats_kernel<T, V><<<grid, kThreads, 0, st>>>(p);
    if (int e = check_launch("cng_group_norm_fwd: statistics")) return e;
    apply_kernel<T, V><<<grid, kThreads, 0, st>>>(p);
    return check_launch("cng_group_norm_fwd");
  }
  bwd_stats_kernel<T, V><<<grid, kThreads, 0, st>>>(p);
  if (int e = check_launch("cng_group_norm_bwd: sums")) return e;
  bwd_apply_kernel<T, V><<<grid, kThreads, 0, st>>>(p);
  return check_launch("cng_group_norm_bwd");
}

template <typename T>
static int dispatch_v(Params& p, bool backward, cudaStream_t st) {
  const int cpg = p.C / p.G;
  constexpr int vmax = Vec<T>::n;
  if (vmax >= 8 && cpg % 8 == 0) return launch<T, (vmax >= 8 ? 8 : vmax)>(p, backward, st);
  if (cpg % 4 == 0) return launch<T, 4>(p, backward, st);
  if (cpg % 2 == 0) return launch<T, 2>(p, backward, st);
  return launch<T, 1>(p, backward, st);
}

static int run(Params& p, int dtype, bool backward, cudaStream_t st) {
  if (dtype == 0) return dispatch_v<float>(p, backward, st);
  if (dtype == 1) return dispatch_v<__half>(p, backward, st);
  return dispatch_v<__nv_bfloat16>(p, backward, st);
}

static int check(const Params& p, int dtype, const char* who) {
  CNG_REQUIRE(dtype >= 0 && dtype <= 2, CNG_ERR_INVALID_ARGUMENT, "%s: dtype %d (0 fp32, 1 fp16, 2 bf16)", who, dtype);
  CNG_REQUIRE(p.N >= 0 && p.S >= 1 && p.C >= 1 && p.G >= 1 && p.C % p.G == 0, CNG_ERR_INVALID_ARGUMENT, "%s: bad shape N=%d S=%lld C=%d G=%d", who, p.N,
              p.S, p.C, p.G);
  CNG_REQUIRE(p.C <= kMaxC && p.G <= 64 && p.N <= 65535, CNG_ERR_UNSUPPORTED, "%s: C=%d (<= %d), G=%d (<= 64), N=%d", who, p.C, kMaxC, p.G, p.N);
  const int cpg = p.C / p.G;
  const int v = (dtype != 0 && cpg % 8 == 0) ? 8 : (cpg % 4 == 0 ? 4 : (cpg % 2 == 0 ? 2 : 1));      // elements per access (dispatch_v)
  CNG_REQUIRE(p.C / v <= kThreads, CNG_ERR_UNSUPPORTED, "%s: C=%d with %d channels per group needs %d threads per position (> %d)", who, p.C, cpg,
              p.C / v, kThreads);
  return CNG_OK;
}

}  // namespace gn
}  // namespace cng

extern "C" {

int cng_group_norm_fwd(const void* x, int dtype, int N, long long S, int C, int G, const float* gamma, const float* beta, float eps, void* y,
                       float* mean, float* rstd, double* sums_scratch, cng_stream_t stream) {
  using namespace cng;
  gn::Params p{};
  p.x = x; p.y = y; p.S = S; p.N = N; p.C = C; p.G = G; p.gamma = gamma; p.beta = beta; p.eps = eps; p.mean = mean; p.rstd = rstd; p.sums = sums_scratch;
  if (int e = gn::check(p, dtype, "group_norm_fwd")) return e;
  if (N == 0) return CNG_OK;
  CNG_REQUIRE(x && y && mean && rstd && sums_scratch, CNG_ERR_INVALID_ARGUMENT, "group_norm_fwd: NULL pointer");
  CNG_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, CNG_ERR_INVALID_ARGUMENT, "group_norm_fwd: x / y not 16-byte aligned");
  if (int e = cng_device_check()) return e;
  cudaStream_t st = as_stream(stream);
  if (cudaMemsetAsync(sums_scratch, 0, static_cast<size_t>(N) * G * 2 * sizeof(double), st) != cudaSuccess)
    return fail(CNG_ERR_INVALID_ARGUMENT, "group_norm_fwd: cannot clear the scratch buffer");
  return gn::run(p, dtype, false, st);
}

int cng_group_norm_bwd(const void* dy, const void* x, int dtype, int N, long long S, int C, int G, const float* gamma, const float* mean,
                       const float* rstd, void* dx, float* ds_scratch, float* db_scratch, cng_stream_t stream) {
  using namespace cng;
  gn::Params p{};
  p.x = x; p.dy = dy; p.y = dx; p.S = S; p.N = N; p.C = C; p.G = G; p.gamma = gamma; p.mean = const_cast<float*>(mean); p.rstd = const_cast<float*>(rstd);
  p.ds = ds_scratch; p.db = db_scratch;
  if (int e = gn::check(p, dtype, "group_norm_bwd")) return e;
  if (N == 0) return CNG_OK;
  CNG_REQUIRE(x && dy && dx && mean && rstd && ds_scratch && db_scratch, CNG_ERR_INVALID_ARGUMENT, "group_norm_bwd: NULL pointer");
  CNG_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0, CNG_ERR_INVALID_ARGUMENT,
              "group_norm_bwd: x / dy / dx not 16-byte aligned");
  if (int e = cng_device_check()) return e;
  cudaStream_t st = as_stream(stream);
  if (cudaMemsetAsync(ds_scratch, 0, static_cast<size_t>(N) * C * sizeof(float), st) != cudaSuccess ||
      cudaMemsetAsync(db_scratch, 0, static_cast<size_t>(N) * C * sizeof(float), st) != cudaSuccess)
    return fail(CNG_ERR_INVALID_ARGUMENT, "group_norm_bwd: cannot clear the scratch buffers");
  return gn::run(p, dtype, true, st);
}

}  // extern "C"
