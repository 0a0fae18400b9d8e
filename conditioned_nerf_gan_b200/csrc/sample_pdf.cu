// K4: inverse-CDF importance resampling, one warp per ray, CDF + binary search in shared memory.
//
// Replaces sample_pdf (generators/volumetric_rendering.py:297-342, det=False) and, in the
// FROM_COARSE instantiation, its call site (generators/generators.py:123-136: midpoints of the
// coarse distances as bins, weights[1:-1] + 1e-5 as weights).
//
// Bit-exactness contract (oracle/nerf_path.py::resample_pdf): the weight sum and the CDF are
// accumulated in float64 and rounded to fp32 per element -- which is what torch.cumsum does for
// fp32 on CPU -- so the searchsorted indices are reproducible bit for bit.  The float64 sums are
// exact for normalised weights (see the kernel), so a lane-blocked warp scan replaces the
// sequential loop; the divide, the search and the lerp run in parallel with IEEE fp32 ops in
// the reference's order (no FMA contraction: every op is an explicit __f*_rn intrinsic).
// Algorithmic bytes per ray: 4*((M+1) + M + K) read + 4*K written (+8*K if indices are emitted).
#include <stdlib.h>

#include "cng_common.cuh"

namespace cng {

struct PdfParams {
  const float* bins;      // [n, M+1]            (FROM_COARSE: t_coarse [n, S])
  const float* weights;   // [n, M]              (FROM_COARSE: raw coarse weights [n, S])
  const float* u;         // [n, K]
  long long n;
  int M, K, S;
  float eps;
  float* samples;         // [n, K]
  int64_t* inds;          // [n, K] or NULL
};

constexpr int kPdfWarps = 8;

template <bool FROM_COARSE>
__global__ void __launch_bounds__(kPdfWarps * 32) sample_pdf_kernel(PdfParams p) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long ray = static_cast<long long>(blockIdx.x) * kPdfWarps + warp;
  if (ray >= p.n) return;
  const int M = p.M;
  const int stride = 2 * (M + 1) + 2;           // cdf[M+1], bins[M+1]
  float* cdf = smem + static_cast<size_t>(warp) * stride;
  float* bins = cdf + (M + 1);

  // ---- load bins and weights (+eps) -------------------------------------------------------
  if (FROM_COARSE) {
    const float* tc = p.bins + ray * p.S;
    const float* wc = p.weights + ray * p.S;
    for (int j = lane; j <= M; j += 32)           // z_vals_mid = 0.5 * (z[:-1] + z[1:])   (:126)
      bins[j] = __fmul_rn(0.5f, __fadd_rn(__ldg(tc + j), __ldg(tc + j + 1)));
    for (int j = lane; j < M; j += 32)            // (weights + 1e-5)[1:-1], then + eps     (:124, :311)
      cdf[j + 1] = __fadd_rn(__fadd_rn(__ldg(wc + j + 1), 1e-5f), p.eps);
  } else {
    for (int j = lane; j <= M; j += 32) bins[j] = __ldg(p.bins + ray * (M + 1) + j);
    for (int j = lane; j < M; j += 32) cdf[j + 1] = __fadd_rn(__ldg(p.weights + ray * M + j), p.eps);
  }
  __syncwarp();
  // ---- total and CDF in float64 ------------------------------------------------------------------
  // The oracle (and torch.cumsum on CPU) accumulates fp32 values sequentially in float64.  Every pdf value
  // is an fp32 number in [~1e-8, 1] (weights carry +eps), so every partial sum is a multiple of 2^-50 that
  // is <= ~1: it fits the 53-bit float64 significand, float64 addition is EXACT here, and any association
  // order -- the lane-blocked scan below included -- produces the same bits as the sequential loop.
  const int per = ((M + 31) >> 5) | 1;            // contiguous elements per lane; odd, so the lanes hit distinct banks
  const int j0 = 1 + lane * per, j1 = min(M + 1, j0 + per);
  double part = 0.0;
  for (int j = j0; j < j1; ++j) part += static_cast<double>(cdf[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  const float total = static_cast<float>(part);
  double run = 0.0;
  for (int j = j0; j < j1; ++j) {
    const float pdf = __fdiv_rn(cdf[j], total);
    cdf[j] = pdf;
    run += static_cast<double>(pdf);
  }
  double incl = run;                              // inclusive scan of the lane totals
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  double acc = incl - run;                        // sum of all earlier lanes
  for (int j = j0; j < j1; ++j) {
    acc += static_cast<double>(cdf[j]);
    cdf[j] = static_cast<float>(acc);
  }
  if (lane == 0) cdf[0] = 0.f;
  __syncwarp();
  // ---- searchsorted(cdf, u, right=False) + lerp ---------------------------------------------
  for (int k = lane; k < p.K; k += 32) {
    const float u = __ldg(p.u + ray * p.K + k);
    int lo = 0, hi = M + 1;                       // first i in [0, M+1] with cdf[i] >= u
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] < u) lo = mid + 1; else hi = mid;
    }
    const int ind = lo;
    const int below = max(ind - 1, 0);
    const int above = min(ind, M);
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = bins[below], b1 = bins[above];
    float denom = __fsub_rn(c1, c0);
    if (denom < p.eps) denom = 1.f;
    const float frac = __fdiv_rn(__fsub_rn(u, c0), denom);
    p.samples[ray * p.K + k] = __fadd_rn(b0, __fmul_rn(frac, __fsub_rn(b1, b0)));
    if (p.inds) p.inds[ray * p.K + k] = ind;
  }
}

// The same arithmetic with the ray's weights / pdf values in REGISTERS (PER consecutive elements per lane, PER = ceil(M / 32) <= 4;
// measured on B200, 1M rays: 64 samples 0.650 -> 0.537 ms, 128 samples 1.046 -> 0.905 ms, 256 samples slower with 8 per lane)
// and a branch-free search: the shared-memory round trips of the scan and the data-dependent loop of the binary search were
// what kept the kernel at 25-31 % of the HBM roofline (issue-bound, profiles/r1g_sample_pdf_kernel_raw.txt).  Every
// floating-point operation and its order are those of the kernel above, so the results are bit-identical.
template <bool FROM_COARSE, int PER, int P2>
__global__ void __launch_bounds__(kPdfWarps * 32) sample_pdf_reg_kernel(PdfParams p) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long ray = static_cast<long long>(blockIdx.x) * kPdfWarps + warp;
  if (ray >= p.n) return;
  const int M = p.M;                                                  // M + 2 <= P2 <= 64 * PER
  float* cdf = smem + static_cast<size_t>(warp) * (P2 + M + 1);      // cdf[0 .. P2): entries beyond M are +inf
  float* bins = cdf + P2;                                             // bins[0 .. M]
  const int j0 = lane * PER;
  const float eps = p.eps;
  float w[PER];
  // (loops over the ray's elements are unrolled with compile-time trip counts: the run-time ones cost ~100 instructions of loop
  //  control and address arithmetic per ray, a fifth of the kernel, profiles/r2h_c5_sass_counts.txt)
  if (FROM_COARSE) {
    const float* tc = p.bins + ray * p.S;
    const float* wc = p.weights + ray * p.S;
#pragma unroll
    for (int i = 0; i <= PER; ++i) {
      const int j = lane + 32 * i;
      if (j <= M) bins[j] = __fmul_rn(0.5f, __fadd_rn(__ldg(tc + j), __ldg(tc + j + 1)));
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) w[i] = (j0 + i < M) ? __fadd_rn(__fadd_rn(__ldg(wc + j0 + i + 1), 1e-5f), eps) : 0.f;
  } else {
    const float* bn = p.bins + ray * (M + 1);
    const float* wn = p.weights + ray * M;
#pragma unroll
    for (int i = 0; i <= PER; ++i) {
      const int j = lane + 32 * i;
      if (j <= M) bins[j] = __ldg(bn + j);
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) w[i] = (j0 + i < M) ? __fadd_rn(__ldg(wn + j0 + i), eps) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < P2 / 32; ++i) {
    const int j = lane + 32 * i;
    if (j > M) cdf[j] = __int_as_float(0x7f800000);
  }
  double part = 0.0;                               // exact in float64 (see the kernel above): any association order gives the same bits
#pragma unroll
  for (int i = 0; i < PER; ++i) part += static_cast<double>(w[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  const float total = static_cast<float>(part);
  double run = 0.0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    w[i] = (j0 + i < M) ? __fdiv_rn(w[i], total) : 0.f;
    run += static_cast<double>(w[i]);
  }
  double incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  double acc = incl - run;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    acc += static_cast<double>(w[i]);
    if (j0 + i < M) cdf[j0 + i + 1] = static_cast<float>(acc);
  }
  if (lane == 0) cdf[0] = 0.f;
  __syncwarp();
  // searchsorted(cdf, u, right=False) = number of cdf[0..M] that are < u; the +inf padding never counts
  const int K = p.K;
  const float* u_ray = p.u + ray * K;
  float* s_ray = p.samples + ray * K;
  const bool has_inds = p.inds != nullptr;
  int64_t* i_ray = has_inds ? p.inds + ray * K : nullptr;
  for (int k = lane; k < K; k += 32) {
    const float u = __ldg(u_ray + k);
    int pos = 0;
#pragma unroll
    for (int step = P2 >> 1; step > 0; step >>= 1) pos += (cdf[pos + step - 1] < u) ? step : 0;
    const int ind = pos;
    const int below = max(ind - 1, 0);
    const int above = min(ind, M);
    const float c0 = cdf[below], c1 = cdf[above];
    const float b0 = bins[below], b1 = bins[above];
    float denom = __fsub_rn(c1, c0);
    if (denom < eps) denom = 1.f;
    const float frac = __fdiv_rn(__fsub_rn(u, c0), denom);
    s_ray[k] = __fadd_rn(b0, __fmul_rn(frac, __fsub_rn(b1, b0)));
    if (has_inds) i_ray[k] = ind;
  }
}

template <bool FROM_COARSE>
static int launch_pdf(const PdfParams& p, cudaStream_t stream, const char* what) {
  const unsigned grid = static_cast<unsigned>((p.n + kPdfWarps - 1) / kPdfWarps);
  const int per = (p.M + 31) / 32;
  // elements per lane the register kernel is used up to: 8 (129..256 bins) measured slower than the shared-memory kernel on B200
  // (2 Mi rays: 200 samples 3.82 vs 3.14 ms, 256 samples 4.67 vs 4.28 ms, same bits), so the default stays 4; CNG_PDF_MAX_PER=8 for A/B
  static const int max_per = [] { const char* e = getenv("CNG_PDF_MAX_PER"); return e ? atoi(e) : 4; }();
  if (per <= max_per && per <= 8) {
    int P2 = 2;
    while (P2 < p.M + 2) P2 <<= 1;                 // P2 - 1 >= M + 1 searched entries
    if (P2 < 32) P2 = 32;
    const size_t smem = static_cast<size_t>(kPdfWarps) * (P2 + p.M + 1) * sizeof(float);
#define CNG_PDF(PER_, P2_) sample_pdf_reg_kernel<FROM_COARSE, PER_, P2_><<<grid, kPdfWarps * 32, smem, stream>>>(p)
    if (per <= 1) { if (P2 == 32) CNG_PDF(1, 32); else CNG_PDF(1, 64); }
    else if (per <= 2) { if (P2 == 64) CNG_PDF(2, 64); else CNG_PDF(2, 128); }
    else if (per <= 4) { if (P2 == 128) CNG_PDF(4, 128); else CNG_PDF(4, 256); }
    else { if (P2 == 256) CNG_PDF(8, 256); else CNG_PDF(8, 512); }
#undef CNG_PDF
    return check_launch(what);
  }
  const size_t smem = static_cast<size_t>(kPdfWarps) * (2 * (p.M + 1) + 2) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(sample_pdf_kernel<FROM_COARSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  sample_pdf_kernel<FROM_COARSE><<<grid, kPdfWarps * 32, smem, stream>>>(p);
  return check_launch(what);
}

}  // namespace cng

extern "C" {

int cng_sample_pdf(const float* bins, const float* weights, const float* u, long long n, int M, int K, float eps,
                   float* samples, int64_t* inds, cng_stream_t stream) {
  CNG_REQUIRE(n >= 0 && M >= 1 && K >= 1, CNG_ERR_INVALID_ARGUMENT, "sample_pdf: n=%lld M=%d K=%d", n, M, K);
  CNG_REQUIRE(n == 0 || (bins && weights && u && samples), CNG_ERR_INVALID_ARGUMENT, "sample_pdf: NULL pointer");
  CNG_REQUIRE(M <= 2047, CNG_ERR_UNSUPPORTED, "sample_pdf: M=%d > 2047", M);
  if (n == 0) return CNG_OK;
  if (int e = cng_device_check()) return e;
  cng::PdfParams p{bins, weights, u, n, M, K, 0, eps, samples, inds};
  return cng::launch_pdf<false>(p, cng::as_stream(stream), "cng_sample_pdf");
}

int cng_resample_from_coarse(const float* t_coarse, const float* weights, const float* u, long long n, int S,
                             float* t_fine, int64_t* inds, cng_stream_t stream) {
  CNG_REQUIRE(n >= 0 && S >= 3, CNG_ERR_INVALID_ARGUMENT, "resample_from_coarse: n=%lld S=%d (need S >= 3)", n, S);
  CNG_REQUIRE(n == 0 || (t_coarse && weights && u && t_fine), CNG_ERR_INVALID_ARGUMENT, "resample_from_coarse: NULL pointer");
  CNG_REQUIRE(S <= 2049, CNG_ERR_UNSUPPORTED, "resample_from_coarse: S=%d > 2049", S);
  if (n == 0) return CNG_OK;
  if (int e = cng_device_check()) return e;
  // bins = S-1 midpoints, weights = S-2 interior weights, N_importance = S
  cng::PdfParams p{t_coarse, weights, u, n, S - 2, S, S, 1e-5f, t_fine, inds};
  return cng::launch_pdf<true>(p, cng::as_stream(stream), "cng_resample_from_coarse");
}

}  // extern "C"
