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
//   cng_film_sin_apply / cng_film_sin_grad   the elementwise halves of FiLMLayer forward / backward
//                             (siren.py:153-157) around the GEMMs of the activation-recomputing
//                             backward (see generators/autograd.py)
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
  load_and_sort_ray(keys, two ? p.t_fine : nullptr, p.t, ray, S, n, n2, lane);

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

// ---- FiLM + sine, elementwise halves --------------------------------------------------------------
// y = sin(freq * (z + b) + phase) as bf16 (the next GEMM's operand);  z [P, HID] fp32 GEMM output
__global__ void __launch_bounds__(256) film_sin_apply_kernel(const float* __restrict__ z, const float* __restrict__ bias,
                                                              const float* __restrict__ freq, const float* __restrict__ phase,
                                                              long long P, int HID, __nv_bfloat16* __restrict__ y) {
  const long long total4 = P * HID / 4;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total4; e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>((e * 4) % HID);
    const float4 v = __ldg(reinterpret_cast<const float4*>(z) + e);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
    const float4 f = __ldg(reinterpret_cast<const float4*>(freq + c));
    const float4 ph = __ldg(reinterpret_cast<const float4*>(phase + c));
    __nv_bfloat162 lo = __floats2bfloat162_rn(sinf(fmaf(f.x, v.x + bb.x, ph.x)), sinf(fmaf(f.y, v.y + bb.y, ph.y)));
    __nv_bfloat162 hi = __floats2bfloat162_rn(sinf(fmaf(f.z, v.z + bb.z, ph.z)), sinf(fmaf(f.w, v.w + bb.w, ph.w)));
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(y)[e] = o;
  }
}

// du = dy * cos(u), u = freq*(z+b)+phase;  dz = du*freq (bf16 out);  dfreq += sum_p du*(z+b);  dphase += sum_p du
// block = 256 threads = 64 column groups (4 columns each, HID == 256) x 4 row lanes; walks kGradRows rows with
// 8-byte (dy, dz) and 16-byte (z) accesses; the column sums go through shared memory and one atomicAdd per column.
constexpr int kGradRows = 128;
__global__ void __launch_bounds__(256) film_sin_grad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ z,
                                                             const float* __restrict__ bias, const float* __restrict__ freq,
                                                             const float* __restrict__ phase, long long P,
                                                             __nv_bfloat16* __restrict__ dz, float* __restrict__ dfreq,
                                                             float* __restrict__ dphase) {
  __shared__ float red[2][4][256];
  const int cgp = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int c = cgp * 4;
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c));
  const float4 f4 = __ldg(reinterpret_cast<const float4*>(freq + c));
  const float4 p4 = __ldg(reinterpret_cast<const float4*>(phase + c));
  const long long r0 = static_cast<long long>(blockIdx.x) * kGradRows;
  const long long r1 = min(P, r0 + kGradRows);
  float af[4] = {0.f, 0.f, 0.f, 0.f}, ap[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (long long r = r0 + ty; r < r1; r += 4) {
    const float4 zv = __ldg(reinterpret_cast<const float4*>(z + r * 256 + c));
    const uint2 g2 = __ldg(reinterpret_cast<const uint2*>(dy + r * 256 + c));
    const __nv_bfloat162 g01 = *reinterpret_cast<const __nv_bfloat162*>(&g2.x), g23 = *reinterpret_cast<const __nv_bfloat162*>(&g2.y);
    const float g[4] = {__low2float(g01), __high2float(g01), __low2float(g23), __high2float(g23)};
    const float zb[4] = {zv.x + b4.x, zv.y + b4.y, zv.z + b4.z, zv.w + b4.w};
    const float fr[4] = {f4.x, f4.y, f4.z, f4.w}, ph[4] = {p4.x, p4.y, p4.z, p4.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float du = g[k] * cosf(fmaf(fr[k], zb[k], ph[k]));
      af[k] = fmaf(du, zb[k], af[k]);
      ap[k] += du;
      o[k] = du * fr[k];
    }
    __nv_bfloat162 o01 = __floats2bfloat162_rn(o[0], o[1]), o23 = __floats2bfloat162_rn(o[2], o[3]);
    uint2 w2;
    w2.x = *reinterpret_cast<uint32_t*>(&o01);
    w2.y = *reinterpret_cast<uint32_t*>(&o23);
    *reinterpret_cast<uint2*>(dz + r * 256 + c) = w2;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { red[0][ty][c + k] = af[k]; red[1][ty][c + k] = ap[k]; }
  __syncthreads();
  const int col = threadIdx.x;
  atomicAdd(dfreq + col, red[0][0][col] + red[0][1][col] + red[0][2][col] + red[0][3][col]);
  atomicAdd(dphase + col, red[1][0][col] + red[1][1][col] + red[1][2][col] + red[1][3][col]);
}

// dz = dy * g (bf16 x fp16 -> bf16), colsum[c] += sum_p dz[p][c];  g = freq * cos(u) dumped by the training-mode forward.
// Same thread layout as film_sin_grad_kernel.
__global__ void __launch_bounds__(256) film_grad_from_g_kernel(const __nv_bfloat16* __restrict__ dy, const __half* __restrict__ g,
                                                                long long P, __nv_bfloat16* __restrict__ dz, float* __restrict__ colsum) {
  __shared__ float red[4][256];
  const int cgp = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int c = cgp * 4;
  const long long r0 = static_cast<long long>(blockIdx.x) * kGradRows;
  const long long r1 = min(P, r0 + kGradRows);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (long long r = r0 + ty; r < r1; r += 4) {
    const uint2 a2 = __ldg(reinterpret_cast<const uint2*>(dy + r * 256 + c));
    const uint2 b2 = __ldg(reinterpret_cast<const uint2*>(g + r * 256 + c));
    const __nv_bfloat162 a01 = *reinterpret_cast<const __nv_bfloat162*>(&a2.x), a23 = *reinterpret_cast<const __nv_bfloat162*>(&a2.y);
    const float2 b01 = __half22float2(*reinterpret_cast<const __half2*>(&b2.x)), b23 = __half22float2(*reinterpret_cast<const __half2*>(&b2.y));
    const float o0 = __low2float(a01) * b01.x, o1 = __high2float(a01) * b01.y;
    const float o2 = __low2float(a23) * b23.x, o3 = __high2float(a23) * b23.y;
    acc[0] += o0; acc[1] += o1; acc[2] += o2; acc[3] += o3;
    __nv_bfloat162 o01 = __floats2bfloat162_rn(o0, o1), o23 = __floats2bfloat162_rn(o2, o3);
    uint2 w2;
    w2.x = *reinterpret_cast<uint32_t*>(&o01);
    w2.y = *reinterpret_cast<uint32_t*>(&o23);
    *reinterpret_cast<uint2*>(dz + r * 256 + c) = w2;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) red[ty][c + k] = acc[k];
  __syncthreads();
  const int col = threadIdx.x;
  atomicAdd(colsum + col, red[0][col] + red[1][col] + red[2][col] + red[3][col]);
}

}  // namespace cng

extern "C" {

int cng_film_grad_from_g(const void* dy_bf16, const void* g_f16, long long P, int HID, void* dz_bf16, float* colsum, cng_stream_t stream) {
  CNG_REQUIRE(P >= 0, CNG_ERR_INVALID_ARGUMENT, "film_grad_from_g: P=%lld", P);
  CNG_REQUIRE(HID == 256, CNG_ERR_UNSUPPORTED, "film_grad_from_g: HID=%d (only 256 is built)", HID);
  if (P == 0) return CNG_OK;
  CNG_REQUIRE(dy_bf16 && g_f16 && dz_bf16 && colsum, CNG_ERR_INVALID_ARGUMENT, "film_grad_from_g: NULL pointer");
  CNG_REQUIRE((P + cng::kGradRows - 1) / cng::kGradRows < 0x7fffffffLL, CNG_ERR_UNSUPPORTED, "film_grad_from_g: too many rows");
  if (int e = cng_device_check()) return e;
  cng::film_grad_from_g_kernel<<<static_cast<unsigned>((P + cng::kGradRows - 1) / cng::kGradRows), 256, 0, cng::as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy_bf16), static_cast<const __half*>(g_f16), P, static_cast<__nv_bfloat16*>(dz_bf16), colsum);
  return cng::check_launch("cng_film_grad_from_g");
}

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

int cng_film_sin_apply(const float* z, const float* bias, const float* freq, const float* phase, long long P, int HID,
                       void* y_bf16, cng_stream_t stream) {
  CNG_REQUIRE(P >= 0 && HID >= 4 && HID % 4 == 0, CNG_ERR_INVALID_ARGUMENT, "film_sin_apply: P=%lld HID=%d", P, HID);
  if (P == 0) return CNG_OK;
  CNG_REQUIRE(z && bias && freq && phase && y_bf16, CNG_ERR_INVALID_ARGUMENT, "film_sin_apply: NULL pointer");
  if (int e = cng_device_check()) return e;
  const long long total4 = P * HID / 4;
  const unsigned grid = static_cast<unsigned>(min((total4 + 255) / 256, static_cast<long long>(cng::sm_count()) * 16));
  cng::film_sin_apply_kernel<<<grid, 256, 0, cng::as_stream(stream)>>>(z, bias, freq, phase, P, HID, static_cast<__nv_bfloat16*>(y_bf16));
  return cng::check_launch("cng_film_sin_apply");
}

int cng_film_sin_grad(const void* dy_bf16, const float* z, const float* bias, const float* freq, const float* phase, long long P,
                      int HID, void* dz_bf16, float* dfreq, float* dphase, cng_stream_t stream) {
  CNG_REQUIRE(P >= 0, CNG_ERR_INVALID_ARGUMENT, "film_sin_grad: P=%lld", P);
  CNG_REQUIRE(HID == 256, CNG_ERR_UNSUPPORTED, "film_sin_grad: HID=%d (only 256 is built)", HID);
  if (P == 0) return CNG_OK;
  CNG_REQUIRE(dy_bf16 && z && bias && freq && phase && dz_bf16 && dfreq && dphase, CNG_ERR_INVALID_ARGUMENT, "film_sin_grad: NULL pointer");
  CNG_REQUIRE((P + cng::kGradRows - 1) / cng::kGradRows < 0x7fffffffLL, CNG_ERR_UNSUPPORTED, "film_sin_grad: too many rows");
  if (int e = cng_device_check()) return e;
  cng::film_sin_grad_kernel<<<static_cast<unsigned>((P + cng::kGradRows - 1) / cng::kGradRows), 256, 0, cng::as_stream(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy_bf16), z, bias, freq, phase, P, static_cast<__nv_bfloat16*>(dz_bf16), dfreq, dphase);
  return cng::check_launch("cng_film_sin_grad");
}

}  // extern "C"
