// K3: alpha compositing, one warp per ray.
//
// Replaces fancy_integration (generators/volumetric_rendering.py:18-70) and, in the MERGE
// instantiation, also the coarse+fine cat/sort/gather (generators/generators.py:163-167), the
// NCHW permute + `*2-1` (:182-183) and distance2depth (:185-186).
//
// Each lane owns IPL consecutive samples of the (depth-sorted) ray.  Transmittance is the
// exclusive product of (1 - alpha + 1e-10): a lane-local running product combined with a
// multiplicative Kogge-Stone scan across lanes (5 __shfl_up_sync steps), so the S' samples of
// a ray are read exactly once and nothing but the per-ray results goes back to HBM.
// Algorithmic bytes per ray: S'*(16+4) read (+4*S' noise) + 16 written (+4*S' if weights).
#include "cng_common.cuh"
#include "merge_sort.cuh"

namespace cng {

struct CompositeParams {
  const float* rgb_sigma;        // [n_rays, S, 4]   (MERGE: coarse)
  const float* rgb_sigma_fine;   // MERGE only: [n_rays, S, 4]
  const float* t;                // [n_rays, S]      (MERGE: coarse)
  const float* t_fine;           // MERGE only
  const float* noise;            // [n_rays, n] or NULL
  const float* rays_d_cam;       // MERGE only: [R, 3]
  long long n_rays;
  int S;                         // samples per input array
  int n;                         // samples composited per ray (S, or 2S when merging)
  int R;                         // rays per image (MERGE)
  float noise_std;
  int clamp_mode, white_back, last_back;
  float* rgb;                    // [n_rays, 3] or NULL
  float* dist;                   // [n_rays] or NULL
  float* weights;                // [n_rays, n] or NULL
  float* pixels;                 // MERGE: [B, 3, R]
  float* depth;                  // MERGE: [B, R]
  int32_t* order;                // MERGE: [n_rays, n] or NULL
};

constexpr int kWarpsPerBlock = 8;

template <int IPL, bool MERGE, bool CHUNKED>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) composite_kernel(CompositeParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long ray = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp;
  if (ray >= p.n_rays) return;
  const int n = p.n;
  const int S = p.S;
  const bool two = MERGE && p.rgb_sigma_fine != nullptr;

  unsigned long long* keys = nullptr;
  if (MERGE) {
    // concatenation order of the reference: fine first, then coarse (generators.py:163-164); stable sort by t.
    // Samples per list this instantiation can meet: n <= 32 * IPL (not chunked), n <= 512 (chunked).
    constexpr int kMinSlots = CHUNKED ? 4 : (IPL == 4 ? 2 : 1), kMaxSlots = CHUNKED ? 8 : (IPL == 4 ? 2 : 1);
    const int n2 = next_pow2_min32(n);
    keys = reinterpret_cast<unsigned long long*>(smem) + static_cast<size_t>(warp) * merge_smem_words(n, S);
    load_and_sort_ray<kMinSlots, kMaxSlots>(keys, two ? p.t_fine : nullptr, p.t, ray, S, n, n2, lane);
  }
  // per-ray bases, hoisted out of the sample loop (the 64-bit ray * S products were a quarter of its instructions)
  const float4* cbase = reinterpret_cast<const float4*>(p.rgb_sigma) + ray * S;          // coarse, or the only list
  const float4* fbase = two ? reinterpret_cast<const float4*>(p.rgb_sigma_fine) + ray * S : cbase;
  const int s_sel = two ? S : 0;                                                           // source index < s_sel: fine list
  const float* tbase = p.t + ray * S;
  const bool has_noise = p.noise != nullptr;
  const float* nbase = has_noise ? p.noise + ray * n : nullptr;
  const bool has_order = MERGE && p.order != nullptr;
  int32_t* obase = has_order ? p.order + ray * n : nullptr;
  const bool has_weights = p.weights != nullptr;
  float* wbase = has_weights ? p.weights + ray * n : nullptr;
  const bool relu = p.clamp_mode == CNG_CLAMP_RELU;
  const float noise_std = p.noise_std;

  // The ray is processed in chunks of 32 * IPL samples: one chunk for up to 128 samples (CHUNKED = false: no loop), chunks of 128
  // beyond -- IPL is capped at 4 so that the per-lane arrays stay in ~24 registers: with 8 or 16 samples per lane the kernel lost
  // its occupancy and fell to 46 % of the HBM roofline at 256 samples per ray (80 % chunked).  The transmittance at the start of
  // a chunk is carried in t_carry.
  float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, wsum = 0.f, t_carry = 1.f;
  const int n_loop = CHUNKED ? n : 1;
  for (int base = 0; base < n_loop; base += 32 * IPL) {
    const int s0 = base + lane * IPL;
    // the lane's IPL distances and the one after them (the last interval of the lane)
    float tv[IPL + 1];
    int src[IPL];
    if (MERGE) {
      // keys[n .. ] is readable scratch, so the loads need no guard; a place past n gets t = 0 (its weight is 0, and 0 * garbage
      // could be NaN in the depth sum)
#pragma unroll
      for (int i = 0; i <= IPL; ++i) {
        const unsigned long long k = keys[s0 + i];
        tv[i] = (s0 + i < n) ? key_t(k) : 0.f;
        if (i < IPL) src[i] = key_src(k);
      }
    } else {
#pragma unroll
      for (int i = 0; i <= IPL; ++i) tv[i] = (s0 + i < n) ? __ldg(tbase + s0 + i) : 0.f;
    }
    float alpha[IPL], fac[IPL], cr[IPL], cg[IPL], cb[IPL];
    float lane_prod = 1.f;
#pragma unroll
    for (int i = 0; i < IPL; ++i) {
      const int s = s0 + i;
      const bool valid = s < n;
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) {
        if (MERGE) {
          const int e = src[i];
          c = __ldg((e < s_sel ? fbase : cbase - s_sel) + e);
          if (has_order) obase[s] = e;
        } else {
          c = __ldg(cbase + s);
        }
      }
      float sg = c.w;
      if (has_noise && valid) sg = sg + __ldg(nbase + s) * noise_std;
      sg = relu ? fmaxf(sg, 0.f) : (sg > 20.f ? sg : log1pf(expf(sg)));        // F.softplus(beta=1, threshold=20)
      const float delta = (s + 1 < n) ? (tv[i + 1] - tv[i]) : 1e10f;
      const float a = valid ? 1.f - expf(-delta * sg) : 0.f;
      alpha[i] = a;
      fac[i] = valid ? (1.f - a) + 1e-10f : 1.f;
      cr[i] = c.x; cg[i] = c.y; cb[i] = c.z;
      lane_prod *= fac[i];
    }
    // exclusive multiplicative scan of lane_prod across the warp
    float incl = lane_prod;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= up;
    }
    float T = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) T = 1.f;
    if (CHUNKED) {
      T *= t_carry;
      t_carry *= __shfl_sync(0xffffffffu, incl, 31);
    }

    float w[IPL];
#pragma unroll
    for (int i = 0; i < IPL; ++i) {
      w[i] = alpha[i] * T;
      T *= fac[i];
      wsum += w[i];
    }
    if (!CHUNKED || base + 32 * IPL >= n) {     // last chunk: the ray's weight sum is complete
      wsum = warp_sum(wsum);
      if (p.last_back) {
        // weights[:, :, -1] += 1 - weights_sum   (volumetric_rendering.py:54-55)
        const int s_last = n - 1 - base;
        if (lane == s_last / IPL) {
#pragma unroll
          for (int i = 0; i < IPL; ++i)
            if (lane * IPL + i == s_last) w[i] += 1.f - wsum;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < IPL; ++i) {
      ar += w[i] * cr[i]; ag += w[i] * cg[i]; ab += w[i] * cb[i]; ad += w[i] * tv[i];
      if (has_weights && s0 + i < n) wbase[s0 + i] = w[i];
    }
  }
  ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); ad = warp_sum(ad);
  if (p.white_back) { const float bg = 1.f - wsum; ar = ar + bg; ag = ag + bg; ab = ab + bg; }
  if (lane == 0) {
    if (p.rgb) { p.rgb[ray * 3 + 0] = ar; p.rgb[ray * 3 + 1] = ag; p.rgb[ray * 3 + 2] = ab; }
    if (p.dist) p.dist[ray] = ad;
    if (MERGE) {
      const long long b = ray / p.R;
      const int r = static_cast<int>(ray - b * p.R);
      if (p.pixels) {
        float* px = p.pixels + b * 3 * p.R + r;
        px[0] = ar * 2.f - 1.f; px[p.R] = ag * 2.f - 1.f; px[2 * static_cast<size_t>(p.R)] = ab * 2.f - 1.f;
      }
      if (p.depth) p.depth[ray] = __ldg(p.rays_d_cam + 3 * r + 2) * ad;
    }
  }
}

template <bool MERGE>
static int launch_composite(const CompositeParams& p, cudaStream_t stream) {
  const int ipl = (p.n + 31) / 32;                    // more than 128 samples: chunks of 128 (IPL = 4)
  const unsigned grid = static_cast<unsigned>((p.n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const size_t smem = MERGE ? static_cast<size_t>(kWarpsPerBlock) * merge_smem_words(p.n, p.S) * sizeof(unsigned long long) : 0;
#define CNG_LAUNCH(I, CH)                                                                          \
  {                                                                                               \
    if (smem > 48 * 1024)                                                                         \
      cudaFuncSetAttribute(composite_kernel<I, MERGE, CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    composite_kernel<I, MERGE, CH><<<grid, kWarpsPerBlock * 32, smem, stream>>>(p);               \
  }
  if (ipl <= 1) CNG_LAUNCH(1, false)
  else if (ipl <= 2) CNG_LAUNCH(2, false)
  else if (ipl <= 4) CNG_LAUNCH(4, false)
  else CNG_LAUNCH(4, true)
#undef CNG_LAUNCH
  return check_launch(MERGE ? "cng_merge_composite" : "cng_composite_fwd");
}

// a11 on its own: stable (fine-first) order of a ray's 2S distances, and optionally the sorted distances.
__global__ void __launch_bounds__(kWarpsPerBlock * 32) merge_sort_kernel(const float* __restrict__ t_fine, const float* __restrict__ t_coarse,
                                                                         long long n_rays, int S, int32_t* __restrict__ order,
                                                                         float* __restrict__ t_sorted) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long ray = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp;
  if (ray >= n_rays) return;
  const int n = 2 * S, n2 = next_pow2_min32(n);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem) + static_cast<size_t>(warp) * merge_smem_words(n, S);
  load_and_sort_ray<1, 8>(keys, t_fine, t_coarse, ray, S, n, n2, lane);
  for (int s = lane; s < n; s += 32) {
    const unsigned long long k = keys[s];
    if (order) order[ray * n + s] = key_src(k);
    if (t_sorted) t_sorted[ray * n + s] = key_t(k);
  }
}

}  // namespace cng

extern "C" {

int cng_merge_sort(const float* t_fine, const float* t_coarse, long long n_rays, int S, int32_t* order, float* t_sorted,
                   cng_stream_t stream) {
  CNG_REQUIRE(n_rays >= 0 && S >= 1, CNG_ERR_INVALID_ARGUMENT, "merge_sort: n_rays=%lld S=%d", n_rays, S);
  CNG_REQUIRE(2 * S <= 512, CNG_ERR_UNSUPPORTED, "merge_sort: %d samples per ray > 512", 2 * S);
  CNG_REQUIRE(n_rays == 0 || (t_fine && t_coarse && (order || t_sorted)), CNG_ERR_INVALID_ARGUMENT, "merge_sort: NULL pointer");
  if (n_rays == 0) return CNG_OK;
  if (int e = cng_device_check()) return e;
  const unsigned grid = static_cast<unsigned>((n_rays + cng::kWarpsPerBlock - 1) / cng::kWarpsPerBlock);
  const size_t smem = static_cast<size_t>(cng::kWarpsPerBlock) * cng::merge_smem_words(2 * S, S) * sizeof(unsigned long long);
  if (smem > 48 * 1024) cudaFuncSetAttribute(cng::merge_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  cng::merge_sort_kernel<<<grid, cng::kWarpsPerBlock * 32, smem, cng::as_stream(stream)>>>(t_fine, t_coarse, n_rays, S, order, t_sorted);
  return cng::check_launch("cng_merge_sort");
}


int cng_composite_fwd(const float* rgb_sigma, const float* t, const float* noise, long long n_rays, int S,
                      float noise_std, int clamp_mode, int white_back, int last_back, float* rgb, float* dist,
                      float* weights, cng_stream_t stream) {
  CNG_REQUIRE(n_rays >= 0 && S >= 1, CNG_ERR_INVALID_ARGUMENT, "composite_fwd: n_rays=%lld S=%d", n_rays, S);
  CNG_REQUIRE(n_rays == 0 || (rgb_sigma && t), CNG_ERR_INVALID_ARGUMENT, "composite_fwd: NULL input");
  CNG_REQUIRE(S <= 1024, CNG_ERR_UNSUPPORTED, "composite_fwd: S=%d > 1024", S);
  CNG_REQUIRE(clamp_mode == CNG_CLAMP_RELU || clamp_mode == CNG_CLAMP_SOFTPLUS, CNG_ERR_INVALID_ARGUMENT,
              "composite_fwd: Need to choose clamp mode");
  CNG_REQUIRE(n_rays / cng::kWarpsPerBlock < 0x7fffffffLL, CNG_ERR_UNSUPPORTED, "composite_fwd: too many rays");
  if (n_rays == 0) return CNG_OK;
  if (int e = cng_device_check()) return e;
  cng::CompositeParams p{};
  p.rgb_sigma = rgb_sigma; p.t = t; p.noise = (noise_std != 0.f) ? noise : nullptr;
  CNG_REQUIRE(noise_std == 0.f || noise, CNG_ERR_INVALID_ARGUMENT, "composite_fwd: noise_std != 0 needs noise");
  p.n_rays = n_rays; p.S = S; p.n = S; p.R = 1; p.noise_std = noise_std; p.clamp_mode = clamp_mode;
  p.white_back = white_back; p.last_back = last_back; p.rgb = rgb; p.dist = dist; p.weights = weights;
  return cng::launch_composite<false>(p, cng::as_stream(stream));
}

int cng_merge_composite(const float* rgb_sigma_fine, const float* rgb_sigma_coarse, const float* t_fine,
                        const float* t_coarse, const float* noise, const float* rays_d_cam, int B, int R, int S,
                        float noise_std, int clamp_mode, int white_back, int last_back, float* pixels, float* depth,
                        float* rgb, float* dist, int32_t* order, cng_stream_t stream) {
  CNG_REQUIRE(B >= 0 && R >= 1 && S >= 1, CNG_ERR_INVALID_ARGUMENT, "merge_composite: B=%d R=%d S=%d", B, R, S);
  CNG_REQUIRE(B == 0 || (rgb_sigma_coarse && t_coarse && rays_d_cam), CNG_ERR_INVALID_ARGUMENT, "merge_composite: NULL input");
  CNG_REQUIRE((rgb_sigma_fine == nullptr) == (t_fine == nullptr), CNG_ERR_INVALID_ARGUMENT,
              "merge_composite: fine rgb_sigma and fine t must both be given or both be NULL");
  const int n = rgb_sigma_fine ? 2 * S : S;
  CNG_REQUIRE(n <= 512, CNG_ERR_UNSUPPORTED, "merge_composite: %d samples per ray > 512", n);
  CNG_REQUIRE(clamp_mode == CNG_CLAMP_RELU || clamp_mode == CNG_CLAMP_SOFTPLUS, CNG_ERR_INVALID_ARGUMENT,
              "merge_composite: Need to choose clamp mode");
  CNG_REQUIRE(noise_std == 0.f || noise, CNG_ERR_INVALID_ARGUMENT, "merge_composite: noise_std != 0 needs noise");
  if (B == 0) return CNG_OK;
  if (int e = cng_device_check()) return e;
  cng::CompositeParams p{};
  p.rgb_sigma = rgb_sigma_coarse; p.rgb_sigma_fine = rgb_sigma_fine; p.t = t_coarse; p.t_fine = t_fine;
  p.noise = (noise_std != 0.f) ? noise : nullptr; p.rays_d_cam = rays_d_cam;
  p.n_rays = static_cast<long long>(B) * R; p.S = S; p.n = n; p.R = R; p.noise_std = noise_std;
  p.clamp_mode = clamp_mode; p.white_back = white_back; p.last_back = last_back;
  p.rgb = rgb; p.dist = dist; p.weights = nullptr; p.pixels = pixels; p.depth = depth; p.order = order;
  return cng::launch_composite<true>(p, cng::as_stream(stream));
}

}  // extern "C"
