// Backward of the rendering path: the gradients autograd computes for
// generators/generators.py:102-180 when the trainer calls loss.backward() (utils.py:711).
//
//   cng_merge_composite_bwd   d(pixels, depth) -> d(rgb_sigma_fine, rgb_sigma_coarse)
//                             backward of cat/sort/gather (generators.py:163-167), fancy_integration
//                             (volumetric_rendering.py:18-70), `*2-1` and distance2depth (:182-186)
//   cng_scatter_points        d(feat) -> d(volume, NDHWC): backward of F.grid_sample w.r.t. its input
//                             (siren.py:555-571); sample positions carry no gradient (they are built
//                             under torch.no_grad(), generators.py:57,111)
//   cng_volume_from_channels_last   NDHWC -> NCDHW (gradient back into the encoder's layout)
//   (the MLP part -- dgrad chain and weight gradients on tcgen05 -- is film_siren_bwd_tc.cu)
//
// One warp per ray for the compositing backward: the forward quantities (merge order, alpha,
// transmittance) are recomputed, the suffix sum over later samples is a reverse warp scan.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "cng_common.cuh"
#include "merge_sort.cuh"

namespace cng {

constexpr int kBwdWarps = 8;

struct CompositeBwdParams {
  const float* rgb_sigma;        // coarse [n_rays, S, 4]
  const float* rgb_sigma_fine;   // [n_rays, S, 4] or NULL
  const float* t;                // coarse [n_rays, S]
  const float* t_fine;
  const float* noise;            // [n_rays, n] or NULL
  const float* rays_d_cam;       // [R, 3]
  const float* d_pixels;         // [B, 3, R] or NULL
  const float* d_depth;          // [B, R] or NULL
  const float* d_rgb;            // un-formatted gradients (plain fancy_integration backward): [n_rays, 3] or NULL
  const float* d_dist;           // [n_rays] or NULL
  long long n_rays;
  int S, n, R;
  float noise_std;
  int clamp_mode, white_back, last_back;
  float* d_rgb_sigma;            // coarse [n_rays, S, 4]
  float* d_rgb_sigma_fine;       // [n_rays, S, 4] or NULL
};

__device__ __forceinline__ float warp_excl_suffix_sum(float v, int lane) {
  // returns sum over lanes > lane
  float incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float dn = __shfl_down_sync(0xffffffffu, incl, o);
    if (lane + o < 32) incl += dn;
  }
  return incl - v;
}

template <int IPL>
__global__ void __launch_bounds__(kBwdWarps * 32) composite_bwd_kernel(CompositeBwdParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long ray = static_cast<long long>(blockIdx.x) * kBwdWarps + warp;
  if (ray >= p.n_rays) return;
  const int n = p.n, S = p.S;
  const bool two = p.rgb_sigma_fine != nullptr;
  const int n2 = next_pow2_min32(n);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem) + static_cast<size_t>(warp) * merge_smem_words(n, S);
  load_and_sort_ray<1, 8>(keys, two ? p.t_fine : nullptr, p.t, ray, S, n, n2, lane);

  const long long b = ray / p.R;
  const int r = static_cast<int>(ray - b * p.R);
  float g0 = 0.f, g1 = 0.f, g2 = 0.f, gd = 0.f;                 // d rgb (before *2-1), d dist
  if (p.d_pixels) {
    const float* dp = p.d_pixels + b * 3 * p.R + r;
    g0 = 2.f * __ldg(dp); g1 = 2.f * __ldg(dp + p.R); g2 = 2.f * __ldg(dp + 2 * static_cast<size_t>(p.R));
  }
  if (p.d_depth) gd = __ldg(p.d_depth + ray) * __ldg(p.rays_d_cam + 3 * r + 2);
  if (p.d_rgb) { g0 = __ldg(p.d_rgb + ray * 3); g1 = __ldg(p.d_rgb + ray * 3 + 1); g2 = __ldg(p.d_rgb + ray * 3 + 2); }
  if (p.d_dist) gd = __ldg(p.d_dist + ray);
  const float gsum = g0 + g1 + g2;

  float alpha[IPL], fac[IPL], e1[IPL], dl[IPL], pre[IPL], G[IPL];
  float4 col[IPL];
  int src[IPL];
  float lane_prod = 1.f;
#pragma unroll
  for (int i = 0; i < IPL; ++i) {
    const int s = lane * IPL + i;
    alpha[i] = 0.f; fac[i] = 1.f; e1[i] = 1.f; dl[i] = 0.f; pre[i] = 0.f; G[i] = 0.f; src[i] = -1;
    col[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (s < n) {
      const unsigned long long k0 = keys[s];
      const float t0 = key_t(k0);
      const float t1 = (s + 1 < n) ? key_t(keys[s + 1]) : 0.f;
      const int e = key_src(k0);
      src[i] = e;
      const float4* sp = (two && e < S) ? reinterpret_cast<const float4*>(p.rgb_sigma_fine) + ray * S + e
                                        : reinterpret_cast<const float4*>(p.rgb_sigma) + ray * S + (two ? e - S : e);
      const float4 c = __ldg(sp);
      col[i] = c;
      const float delta = (s + 1 < n) ? (t1 - t0) : 1e10f;
      float sg = c.w;
      if (p.noise != nullptr) sg = sg + __ldg(p.noise + ray * n + s) * p.noise_std;
      pre[i] = sg;                                              // clamp input
      const float sc = (p.clamp_mode == CNG_CLAMP_RELU) ? fmaxf(sg, 0.f) : (sg > 20.f ? sg : log1pf(expf(sg)));
      const float ex = expf(-delta * sc);
      e1[i] = ex;
      dl[i] = delta;
      alpha[i] = 1.f - ex;
      fac[i] = (1.f - alpha[i]) + 1e-10f;
      G[i] = g0 * c.x + g1 * c.y + g2 * c.z + gd * t0;
      lane_prod *= fac[i];
    }
  }
  float incl = lane_prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= up;
  }
  float T0 = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T0 = 1.f;
  float T[IPL], w[IPL];
  float wsum = 0.f;
  {
    float Tr = T0;
#pragma unroll
    for (int i = 0; i < IPL; ++i) { T[i] = Tr; w[i] = alpha[i] * Tr; Tr *= fac[i]; wsum += w[i]; }
  }
  wsum = warp_sum(wsum);
  // coefficient of the last (farthest) sample, needed by last_back
  const int s_last = n - 1;
  float G_last = 0.f;
  {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < IPL; ++i)
      if (lane * IPL + i == s_last) v = G[i];
    G_last = warp_sum(v);
  }
  // dL/dw_i
  float dw[IPL];
  float lane_dw_w = 0.f;
#pragma unroll
  for (int i = 0; i < IPL; ++i) {
    dw[i] = G[i] - (p.white_back ? gsum : 0.f) - (p.last_back ? G_last : 0.f);
    if (lane * IPL + i >= n) dw[i] = 0.f;
    lane_dw_w += dw[i] * w[i];
  }
  float suffix = warp_excl_suffix_sum(lane_dw_w, lane);         // sum over later lanes of dw*w
  // walk this lane's samples from the far end
#pragma unroll
  for (int i = IPL - 1; i >= 0; --i) {
    const int s = lane * IPL + i;
    if (s < n) {
      const float dalpha = T[i] * dw[i] - suffix / fac[i];
      float dsc = dalpha * dl[i] * e1[i];                        // d alpha / d sigma' = delta * exp(-delta sigma')
      float dclamp;
      if (p.clamp_mode == CNG_CLAMP_RELU) dclamp = pre[i] > 0.f ? 1.f : 0.f;
      else dclamp = pre[i] > 20.f ? 1.f : 1.f / (1.f + expf(-pre[i]));
      if (dclamp == 0.f) dsc = 0.f;                              // keeps 1e10 * 0 out of the product
      float wi = w[i];
      if (p.last_back && s == s_last) wi += 1.f - wsum;
      float4 o;
      o.x = wi * g0; o.y = wi * g1; o.z = wi * g2; o.w = dsc * dclamp;
      const int e = src[i];
      float4* dst = (two && e < S) ? reinterpret_cast<float4*>(p.d_rgb_sigma_fine) + ray * S + e
                                   : reinterpret_cast<float4*>(p.d_rgb_sigma) + ray * S + (two ? e - S : e);
      *dst = o;
    }
    suffix += dw[i] * w[i];
  }
}

template <int IPL>
static void launch_cbwd(const CompositeBwdParams& p, cudaStream_t stream) {
  const unsigned grid = static_cast<unsigned>((p.n_rays + kBwdWarps - 1) / kBwdWarps);
  const size_t smem = static_cast<size_t>(kBwdWarps) * merge_smem_words(p.n, p.S) * sizeof(unsigned long long);
  if (smem > 48 * 1024) cudaFuncSetAttribute(composite_bwd_kernel<IPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  composite_bwd_kernel<IPL><<<grid, kBwdWarps * 32, smem, stream>>>(p);
}

// ---- trilinear scatter-add ------------------------------------------------------------------------
__device__ __forceinline__ void axis_index_bwd(float p, int size, int& i0, float& w_lo, float& w_hi) {
  const float g = __fdiv_rn(p, 0.6f);
  float i = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), static_cast<float>(size)), 1.f), 2.f);
  i = fminf(static_cast<float>(size - 1), fmaxf(i, 0.f));
  const float f = floorf(i);
  i0 = static_cast<int>(f);
  w_lo = __fsub_rn(__fadd_rn(f, 1.f), i);
  w_hi = __fsub_rn(i, f);
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// one point per 8 lanes, each lane owns 4 channels per step; 8 corners -> 8 vector reductions per lane
__global__ void __launch_bounds__(256) scatter_points_kernel(float* __restrict__ dvol_all, int C4, int D, int H, int W,
                                                              const float* __restrict__ points, long long N, long long total,
                                                              const float4* __restrict__ dfeat) {
  const int sub = threadIdx.x & 7;
  const long long i = static_cast<long long>(blockIdx.x) * 32 + (threadIdx.x >> 3);
  if (i >= total) return;
  const long long b = i / N;
  float* dvol = dvol_all + static_cast<size_t>(b) * D * H * W * C4 * 4;
  const float px = __ldg(points + 3 * i), py = __ldg(points + 3 * i + 1), pz = __ldg(points + 3 * i + 2);
  int x0, y0, z0;
  float xl, xh, yl, yh, zl, zh;
  axis_index_bwd(px, W, x0, xl, xh);
  axis_index_bwd(py, H, y0, yl, yh);
  axis_index_bwd(pz, D, z0, zl, zh);
  const bool xin = x0 + 1 <= W - 1, yin = y0 + 1 <= H - 1, zin = z0 + 1 <= D - 1;
  const size_t sx = static_cast<size_t>(C4) * 4, sy = static_cast<size_t>(W) * sx, sz = static_cast<size_t>(H) * sy;
  for (int cg = sub; cg < C4; cg += 8) {
    const float4 g = __ldg(dfeat + i * C4 + cg);
    float* base = dvol + z0 * sz + y0 * sy + x0 * sx + cg * 4;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int dx = k & 1, dy = (k >> 1) & 1, dz = k >> 2;
      if ((dx && !xin) || (dy && !yin) || (dz && !zin)) continue;    // ATen skips out-of-range corners
      const float wgt = (dx ? xh : xl) * (dy ? yh : yl) * (dz ? zh : zl);
      if (wgt == 0.f) continue;
      red_add_v4(base + dz * sz + dy * sy + dx * sx, g.x * wgt, g.y * wgt, g.z * wgt, g.w * wgt);
    }
  }
}

// NDHWC -> NCDHW (reverse of channels_last_kernel)
__global__ void __launch_bounds__(256) channels_first_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, long long vox) {
  extern __shared__ float tile[];   // [C][33]
  const long long v0 = static_cast<long long>(blockIdx.x) * 32;
  const int b = blockIdx.y;
  const float* s = src + static_cast<size_t>(b) * C * vox;
  float* d = dst + static_cast<size_t>(b) * C * vox;
  const int nv = static_cast<int>(min(32LL, vox - v0));
  for (int e = threadIdx.x; e < nv * C; e += 256) {
    const int vv = e / C, c = e - vv * C;
    tile[c * 33 + vv] = __ldg(s + (v0 + vv) * C + c);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int c = w; c < C; c += 8)
    if (lane < nv) d[static_cast<size_t>(c) * vox + v0 + lane] = tile[c * 33 + lane];
}

}  // namespace cng

extern "C" {

int cng_merge_composite_bwd(const float* rgb_sigma_fine, const float* rgb_sigma_coarse, const float* t_fine,
                            const float* t_coarse, const float* noise, const float* rays_d_cam, const float* d_pixels,
                            const float* d_depth, int B, int R, int S, float noise_std, int clamp_mode, int white_back,
                            int last_back, float* d_rgb_sigma_fine, float* d_rgb_sigma_coarse, cng_stream_t stream) {
  CNG_REQUIRE(B >= 0 && R >= 1 && S >= 1, CNG_ERR_INVALID_ARGUMENT, "merge_composite_bwd: B=%d R=%d S=%d", B, R, S);
  if (B == 0) return CNG_OK;
  CNG_REQUIRE(rgb_sigma_coarse && t_coarse && rays_d_cam && d_rgb_sigma_coarse, CNG_ERR_INVALID_ARGUMENT, "merge_composite_bwd: NULL input");
  const bool two = rgb_sigma_fine != nullptr;
  CNG_REQUIRE(two == (t_fine != nullptr) && two == (d_rgb_sigma_fine != nullptr), CNG_ERR_INVALID_ARGUMENT,
              "merge_composite_bwd: fine rgb_sigma, t and gradient must all be given or all be NULL");
  const int n = two ? 2 * S : S;
  CNG_REQUIRE(n <= 512, CNG_ERR_UNSUPPORTED, "merge_composite_bwd: %d samples per ray > 512", n);
  CNG_REQUIRE(clamp_mode == CNG_CLAMP_RELU || clamp_mode == CNG_CLAMP_SOFTPLUS, CNG_ERR_INVALID_ARGUMENT,
              "merge_composite_bwd: Need to choose clamp mode");
  CNG_REQUIRE(noise_std == 0.f || noise, CNG_ERR_INVALID_ARGUMENT, "merge_composite_bwd: noise_std != 0 needs noise");
  if (int e = cng_device_check()) return e;
  cng::CompositeBwdParams p{};
  p.rgb_sigma = rgb_sigma_coarse; p.rgb_sigma_fine = rgb_sigma_fine; p.t = t_coarse; p.t_fine = t_fine;
  p.noise = (noise_std != 0.f) ? noise : nullptr; p.rays_d_cam = rays_d_cam; p.d_pixels = d_pixels; p.d_depth = d_depth;
  p.n_rays = static_cast<long long>(B) * R; p.S = S; p.n = n; p.R = R; p.noise_std = noise_std; p.clamp_mode = clamp_mode;
  p.white_back = white_back; p.last_back = last_back; p.d_rgb_sigma = d_rgb_sigma_coarse; p.d_rgb_sigma_fine = d_rgb_sigma_fine;
  const int ipl = (n + 31) / 32;
  cudaStream_t st = cng::as_stream(stream);
  if (ipl <= 1) cng::launch_cbwd<1>(p, st);
  else if (ipl <= 2) cng::launch_cbwd<2>(p, st);
  else if (ipl <= 4) cng::launch_cbwd<4>(p, st);
  else if (ipl <= 8) cng::launch_cbwd<8>(p, st);
  else cng::launch_cbwd<16>(p, st);
  return cng::check_launch("cng_merge_composite_bwd");
}

int cng_composite_bwd(const float* rgb_sigma, const float* t, const float* noise, const float* d_rgb, const float* d_dist, long long n_rays,
                      int S, float noise_std, int clamp_mode, int white_back, int last_back, float* d_rgb_sigma, cng_stream_t stream) {
  CNG_REQUIRE(n_rays >= 0 && S >= 1, CNG_ERR_INVALID_ARGUMENT, "composite_bwd: n_rays=%lld S=%d", n_rays, S);
  CNG_REQUIRE(S <= 512, CNG_ERR_UNSUPPORTED, "composite_bwd: S=%d > 512", S);
  if (n_rays == 0) return CNG_OK;
  CNG_REQUIRE(rgb_sigma && t && d_rgb_sigma, CNG_ERR_INVALID_ARGUMENT, "composite_bwd: NULL pointer");
  CNG_REQUIRE(clamp_mode == CNG_CLAMP_RELU || clamp_mode == CNG_CLAMP_SOFTPLUS, CNG_ERR_INVALID_ARGUMENT, "composite_bwd: Need to choose clamp mode");
  CNG_REQUIRE(noise_std == 0.f || noise, CNG_ERR_INVALID_ARGUMENT, "composite_bwd: noise_std != 0 needs noise");
  CNG_REQUIRE(n_rays / cng::kBwdWarps < 0x7fffffffLL, CNG_ERR_UNSUPPORTED, "composite_bwd: too many rays");
  if (int e = cng_device_check()) return e;
  cng::CompositeBwdParams p{};
  p.rgb_sigma = rgb_sigma; p.t = t; p.noise = (noise_std != 0.f) ? noise : nullptr; p.d_rgb = d_rgb; p.d_dist = d_dist;
  p.n_rays = n_rays; p.S = S; p.n = S; p.R = 1; p.noise_std = noise_std; p.clamp_mode = clamp_mode;
  p.white_back = white_back; p.last_back = last_back; p.d_rgb_sigma = d_rgb_sigma;
  const int ipl = (S + 31) / 32;
  cudaStream_t st = cng::as_stream(stream);
  if (ipl <= 1) cng::launch_cbwd<1>(p, st);
  else if (ipl <= 2) cng::launch_cbwd<2>(p, st);
  else if (ipl <= 4) cng::launch_cbwd<4>(p, st);
  else if (ipl <= 8) cng::launch_cbwd<8>(p, st);
  else cng::launch_cbwd<16>(p, st);
  return cng::check_launch("cng_composite_bwd");
}

int cng_scatter_points(float* dvol_ndhwc, int B, int C, int D, int H, int W, const float* points, long long N,
                       const float* dfeat, cng_stream_t stream) {
  CNG_REQUIRE(B >= 0 && C >= 1 && D >= 1 && H >= 1 && W >= 1 && N >= 0, CNG_ERR_INVALID_ARGUMENT, "scatter_points: bad shape");
  CNG_REQUIRE(C % 4 == 0 && C <= 128, CNG_ERR_UNSUPPORTED, "scatter_points: C=%d (need C %% 4 == 0 and C <= 128)", C);
  const long long total = static_cast<long long>(B) * N;
  if (total == 0) return CNG_OK;
  CNG_REQUIRE(dvol_ndhwc && points && dfeat, CNG_ERR_INVALID_ARGUMENT, "scatter_points: NULL pointer");
  CNG_REQUIRE(((reinterpret_cast<uintptr_t>(dvol_ndhwc) | reinterpret_cast<uintptr_t>(dfeat)) & 15) == 0, CNG_ERR_INVALID_ARGUMENT,
              "scatter_points: buffers not 16-byte aligned");
  CNG_REQUIRE((total + 31) / 32 < 0x7fffffffLL, CNG_ERR_UNSUPPORTED, "scatter_points: too many points");
  if (int e = cng_device_check()) return e;
  cng::scatter_points_kernel<<<static_cast<unsigned>((total + 31) / 32), 256, 0, cng::as_stream(stream)>>>(
      dvol_ndhwc, C / 4, D, H, W, points, N, total, reinterpret_cast<const float4*>(dfeat));
  return cng::check_launch("cng_scatter_points");
}

int cng_volume_from_channels_last(const float* vol_ndhwc, float* vol_ncdhw, int B, int C, int D, int H, int W, cng_stream_t stream) {
  CNG_REQUIRE(B >= 0 && C >= 1 && D >= 1 && H >= 1 && W >= 1, CNG_ERR_INVALID_ARGUMENT, "volume_from_channels_last: bad shape");
  CNG_REQUIRE(C <= 256 && B <= 65535, CNG_ERR_UNSUPPORTED, "volume_from_channels_last: C=%d B=%d", C, B);
  if (B == 0) return CNG_OK;
  CNG_REQUIRE(vol_ndhwc && vol_ncdhw, CNG_ERR_INVALID_ARGUMENT, "volume_from_channels_last: NULL pointer");
  if (int e = cng_device_check()) return e;
  const long long vox = static_cast<long long>(D) * H * W;
  dim3 grid(static_cast<unsigned>((vox + 31) / 32), B);
  cng::channels_first_kernel<<<grid, 256, static_cast<size_t>(C) * 33 * sizeof(float), cng::as_stream(stream)>>>(vol_ndhwc, vol_ncdhw, C, vox);
  return cng::check_launch("cng_volume_from_channels_last");
}

}  // extern "C"
