// K2 (exact mode): FiLM-SIREN MLP in fp32 FFMA tiles, all layers fused per 64-point tile.
//
// Replaces FiLMLayer.forward x L (generators/siren.py:146-160 inside the loops at :573-577,
// :661-665, :820-824, :1058-1062) and the head nn.Linear(HID,4) + _sigmoid_rgb (:579,
// :1227-1234).  This is the CNG_PREC_FP32 path: same arithmetic as the reference's fp32
// inference (fp32 products, fp32 accumulate, sinf), used where <=1e-5 parity matters; the
// throughput path is film_siren_tc.cu (tcgen05).
//
// A block owns a 64-point tile of one batch item; its activations ping-pong between two
// [64][HID+4] fp32 shared-memory buffers and never go to HBM.  Weights stream from L2 in
// [16][HID] k-slabs (transposed on the way in so a warp reads 32 consecutive outputs).
// Each thread accumulates 8 points x 8 outputs.
#include "cng_common.cuh"

namespace cng {

constexpr int kSimtTM = 64;
constexpr int kSimtKC = 16;
constexpr int kSimtMaxL = 16;

struct SimtParams {
  const float* feat;       // [B, N, C]
  long long N;
  int B, C, HID, L;
  const float* w[kSimtMaxL];
  const float* b[kSimtMaxL];
  const float* freq;       // [B, L*HID]
  const float* phase;      // [B, L*HID]
  const float* final_w;    // [4, HID]
  const float* final_b;    // [4]
  int sigmoid_rgb;
  float* out;              // [B, N, 4]
  long long tiles_per_item;
  unsigned res_save_mask, res_add_mask;    // residual blocks (generators/siren.py:218-230), see film_siren_tc_common.cuh
};

// HID is fixed at 256 (one output per (lane, j) pair: 32 lanes x 8).
__global__ void __launch_bounds__(256, 1) film_siren_simt_kernel(SimtParams p) {
  constexpr int HID = 256;
  constexpr int LD = HID + 4;
  extern __shared__ __align__(16) float smem_f[];
  float* act0 = smem_f;                         // [64][LD]
  float* act1 = act0 + kSimtTM * LD;            // [64][LD]
  float* wc = act1 + kSimtTM * LD;              // [16][HID]
  float* wf = wc + kSimtKC * HID;               // [HID][4]
  float* res = wf + HID * 4;                    // [64][LD], only allocated when a residual mask is set

  const int tid = threadIdx.x, lane = tid & 31, ty = tid >> 5;
  const long long item = blockIdx.x / p.tiles_per_item;
  const long long tile = blockIdx.x - item * p.tiles_per_item;
  const long long n0 = tile * kSimtTM;
  const int rows = static_cast<int>(min(static_cast<long long>(kSimtTM), p.N - n0));
  const int C = p.C;
  const int K0 = (C + kSimtKC - 1) / kSimtKC * kSimtKC;

  // features -> act0 (zero padded to a multiple of 16 columns and to 64 rows)
  const float* f = p.feat + (item * p.N + n0) * C;
  for (int e = tid; e < kSimtTM * K0; e += 256) {
    const int r = e / K0, c = e - r * K0;
    act0[r * LD + c] = (r < rows && c < C) ? __ldg(f + static_cast<size_t>(r) * C + c) : 0.f;
  }
  for (int e = tid; e < HID * 4; e += 256) {
    const int k = e >> 2, j = e & 3;
    wf[e] = __ldg(p.final_w + j * HID + k);
  }
  float* in = act0;
  float* outb = act1;
  for (int l = 0; l < p.L; ++l) {
    const int K = (l == 0) ? C : HID;
    const int Kp = (l == 0) ? K0 : HID;
    const float* W = p.w[l];
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < Kp; k0 += kSimtKC) {
      __syncthreads();   // previous slab consumed / activations of the previous layer complete
      {
        const float* wr = W + static_cast<size_t>(tid) * K + k0;   // output row `tid`
#pragma unroll
        for (int kk = 0; kk < kSimtKC; ++kk) wc[kk * HID + tid] = (k0 + kk < K) ? __ldg(wr + kk) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k4 = 0; k4 < kSimtKC; k4 += 4) {
        float4 x4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x4[i] = *reinterpret_cast<const float4*>(in + (ty * 8 + i) * LD + k0 + k4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          float wv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) wv[j] = wc[(k4 + kk) * HID + lane + 32 * j];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xv = kk == 0 ? x4[i].x : kk == 1 ? x4[i].y : kk == 2 ? x4[i].z : x4[i].w;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xv, wv[j], acc[i][j]);
          }
        }
      }
    }
    // FiLM + sine epilogue: sin(freq * (x W^T + b) + phase), separate mul/add as in eager torch
    const float* fr = p.freq + (item * p.L + l) * HID;
    const float* ph = p.phase + (item * p.L + l) * HID;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = lane + 32 * j;
      const float bias = __ldg(p.b[l] + o), fq = __ldg(fr + o), pq = __ldg(ph + o);
      const bool add = (p.res_add_mask >> l) & 1u, save = (p.res_save_mask >> l) & 1u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float u = __fadd_rn(__fmul_rn(fq, __fadd_rn(acc[i][j], bias)), pq);
        if (add) u = __fadd_rn(res[(ty * 8 + i) * LD + o], u);          // sin(x + fc2(net)), ResSirenBlock.forward
        const float y = sinf(u);
        outb[(ty * 8 + i) * LD + o] = y;
        if (save) res[(ty * 8 + i) * LD + o] = y;                       // same thread reads it back: no barrier needed
      }
    }
    float* tmp = in; in = outb; outb = tmp;
  }
  __syncthreads();
  // head: 64 points x 4 outputs, one per thread
  {
    const int pt = tid >> 2, j = tid & 3;
    float a = 0.f;
    const float* x = in + pt * LD;
#pragma unroll 8
    for (int k = 0; k < HID; ++k) a = fmaf(x[k], wf[k * 4 + j], a);
    a += __ldg(p.final_b + j);
    if (p.sigmoid_rgb && j < 3) a = 1.f / (1.f + expf(-a));
    if (pt < rows) p.out[(item * p.N + n0 + pt) * 4 + j] = a;
  }
}

// a5: FiLM parameters.  freq = 15 * (W g + b)[:half] + 30, phase = (W g + b)[half:]   (generators/siren.py:550-553).
// One warp per output row: lanes stride over z_dim (coalesced weight reads), fixed shuffle-tree reduction, so the result
// for an item does not depend on how many items are in the batch (cuBLAS picks batch-size-dependent kernels, which
// breaks that by an ulp).  Items are processed in chunks of kFilmItems per block so a weight row is read once per chunk.
constexpr int kFilmItems = 8;
__global__ void __launch_bounds__(256) film_parameters_kernel(const float* __restrict__ glob, const float* __restrict__ w,
                                                               const float* __restrict__ bias, int B, int z_dim, int n_out,
                                                               float* __restrict__ freq, float* __restrict__ phase) {
  extern __shared__ float g_s[];                       // [kFilmItems][z_dim]
  const int b0 = blockIdx.y * kFilmItems;
  const int nb = min(kFilmItems, B - b0);
  for (int e = threadIdx.x; e < nb * z_dim; e += blockDim.x) g_s[e] = __ldg(glob + static_cast<size_t>(b0) * z_dim + e);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= n_out) return;
  const float* wr = w + static_cast<size_t>(o) * z_dim;
  float acc[kFilmItems];
#pragma unroll
  for (int i = 0; i < kFilmItems; ++i) acc[i] = 0.f;
  for (int k = lane; k < z_dim; k += 32) {
    const float wv = __ldg(wr + k);
#pragma unroll
    for (int i = 0; i < kFilmItems; ++i)
      if (i < nb) acc[i] = fmaf(g_s[i * z_dim + k], wv, acc[i]);
  }
  const float bv = __ldg(bias + o);
  const int half = n_out / 2;
#pragma unroll
  for (int i = 0; i < kFilmItems; ++i) {
    if (i >= nb) break;
    float v = acc[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    v += bv;
    if (lane == 0) {
      if (o < half) freq[static_cast<size_t>(b0 + i) * half + o] = fmaf(v, 15.f, 30.f);
      else phase[static_cast<size_t>(b0 + i) * half + (o - half)] = v;
    }
  }
}

int film_siren_simt_launch(const float* feat, int B, long long N, int C, int HID, int L, const float* const* w,
                           const float* const* b, const float* freq, const float* phase, const float* final_w,
                           const float* final_b, int sigmoid_rgb, float* out, cudaStream_t stream, unsigned res_save_mask,
                           unsigned res_add_mask) {
  CNG_REQUIRE(HID == 256, CNG_ERR_UNSUPPORTED, "film_siren_fwd(fp32): HID=%d (only 256 is built)", HID);
  CNG_REQUIRE(L >= 1 && L <= kSimtMaxL, CNG_ERR_UNSUPPORTED, "film_siren_fwd(fp32): L=%d", L);
  CNG_REQUIRE(C >= 1 && C <= 256, CNG_ERR_UNSUPPORTED, "film_siren_fwd(fp32): C=%d", C);
  SimtParams p{};
  p.feat = feat; p.N = N; p.B = B; p.C = C; p.HID = HID; p.L = L;
  for (int l = 0; l < L; ++l) { p.w[l] = w[l]; p.b[l] = b[l]; }
  p.freq = freq; p.phase = phase; p.final_w = final_w; p.final_b = final_b; p.sigmoid_rgb = sigmoid_rgb; p.out = out;
  p.res_save_mask = res_save_mask; p.res_add_mask = res_add_mask;
  p.tiles_per_item = (N + kSimtTM - 1) / kSimtTM;
  const long long blocks = p.tiles_per_item * B;
  CNG_REQUIRE(blocks < 0x7fffffffLL, CNG_ERR_UNSUPPORTED, "film_siren_fwd(fp32): too many tiles");
  const size_t smem = ((2 + ((res_save_mask | res_add_mask) ? 1 : 0)) * kSimtTM * (256 + 4) + kSimtKC * 256 + 256 * 4) * sizeof(float);
  cudaFuncSetAttribute(film_siren_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  film_siren_simt_kernel<<<static_cast<unsigned>(blocks), 256, smem, stream>>>(p);
  return check_launch("cng_film_siren_fwd(fp32)");
}

}  // namespace cng

extern "C" int cng_film_parameters(const float* global_feature, const float* map_w, const float* map_b, int B, int z_dim, int n_out,
                                   float* freq, float* phase, cng_stream_t stream) {
  CNG_REQUIRE(B >= 0 && z_dim >= 1 && n_out >= 2 && n_out % 2 == 0, CNG_ERR_INVALID_ARGUMENT, "film_parameters: B=%d z_dim=%d n_out=%d", B, z_dim, n_out);
  CNG_REQUIRE(z_dim <= 1024 && B <= 65535 * 8, CNG_ERR_UNSUPPORTED, "film_parameters: z_dim=%d B=%d", z_dim, B);
  if (B == 0) return CNG_OK;
  CNG_REQUIRE(global_feature && map_w && map_b && freq && phase, CNG_ERR_INVALID_ARGUMENT, "film_parameters: NULL pointer");
  if (int e = cng_device_check()) return e;
  dim3 grid((n_out + 7) / 8, (B + cng::kFilmItems - 1) / cng::kFilmItems);
  cng::film_parameters_kernel<<<grid, 256, static_cast<size_t>(cng::kFilmItems) * z_dim * sizeof(float), cng::as_stream(stream)>>>(
      global_feature, map_w, map_b, B, z_dim, n_out, freq, phase);
  return cng::check_launch("cng_film_parameters");
}
