// cng_render_fwd: the whole forward of ImplicitGenerator3d (generators/generators.py:33-187) behind ONE C-ABI call.
// It sequences the kernels of this library on the caller's stream -- K1 coarse, K2, K3, K4, K1 fine, K2, K3' -- with all
// intermediates in a caller-provided workspace; nothing synchronises.  Random draws are inputs (the reference draws them
// with torch.rand / torch.randn in a fixed order), the volume is NDHWC (cng_volume_to_channels_last),
// freq / phase come from cng_film_parameters.
#include <stdlib.h>

#include "cng_common.cuh"

namespace cng {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// CNG_FUSED_GATHER=1: K1 writes the sample positions only (12 B per point) and K2 looks the features up in its prologue
// (cng_film_siren_fwd_gather); 0: K1 gathers into feat[B,R,S,32] (128 B per point written, then read by K2).  Same bits.
// Measured A/B on B200 at BASELINE configs[1] (profiles/r2c_fused_gather_ab.txt): separate 6.615 ms per step, fused 6.64-6.68 ms.
// The fused form deletes 1.6 GB of HBM traffic and K1's 0.6 ms, but puts the lookup (8 dependent-latency corner loads per lane
// and round, four rounds per tile) on the critical chain of a tile slot, where no other warp of the slot can hide it; K1 as a
// kernel of its own hides the same latency behind 2048 resident threads per SM.  Default: separate.
constexpr int kDefaultFusedGather = 0;
static int g_fused_override = -1;       // debug hook cng_internal_set_fused_gather (tests, A/B tools): -1 = environment / default
static bool fused_gather(int C, int HID, int precision) {
  static const int env_mode = [] {
    const char* e = getenv("CNG_FUSED_GATHER");
    return e ? atoi(e) : kDefaultFusedGather;
  }();
  const int mode = g_fused_override >= 0 ? g_fused_override : env_mode;
  return mode != 0 && C == 32 && HID == 256 && (precision == CNG_PREC_BF16 || precision == CNG_PREC_FP16);
}

struct RenderLayout {
  size_t feat, t_c, rs_c, w_c, t_f, rs_f, mlp, total;
};

static RenderLayout render_layout(int B, long long R, int S, int C, int HID, int L, int hierarchical, int precision) {
  RenderLayout l{};
  const size_t pts = static_cast<size_t>(B) * R * S;
  size_t off = 0;
  l.feat = off; off = align_up(off + pts * C * sizeof(float), 256);
  l.t_c = off; off = align_up(off + pts * sizeof(float), 256);
  l.rs_c = off; off = align_up(off + pts * 4 * sizeof(float), 256);
  if (hierarchical) {
    l.w_c = off; off = align_up(off + pts * sizeof(float), 256);
    l.t_f = off; off = align_up(off + pts * sizeof(float), 256);
    l.rs_f = off; off = align_up(off + pts * 4 * sizeof(float), 256);
  }
  l.mlp = off; off = align_up(off + cng_film_siren_workspace_bytes(B, C, HID, L, precision), 256);
  l.total = off;
  return l;
}

}  // namespace cng

extern "C" {

// Debug hook (not part of the ABI in include/cng_b200.h): 1 / 0 force the fused-gather / separate-gather forward, -1 restores
// CNG_FUSED_GATHER / the built-in default.
CNG_API void cng_internal_set_fused_gather(int mode) { cng::g_fused_override = mode < 0 ? -1 : (mode ? 1 : 0); }

size_t cng_render_workspace_bytes(int B, int img_w, int img_h, int S, int C, int HID, int L, int hierarchical, int precision) {
  if (B <= 0 || img_w <= 0 || img_h <= 0 || S <= 0 || C <= 0 || L <= 0) return 0;
  return cng::render_layout(B, static_cast<long long>(img_w) * img_h, S, C, HID, L, hierarchical, precision).total;
}

int cng_render_fwd(const float* vol_ndhwc, long long vol_item_stride, int B, int C, int D, int H, int W, const float* cam2world,
                   const float* rays_d_cam, const float* t_lin, int img_w, int img_h, int S, int HID, int L,
                   const float* const* layer_w_host, const float* const* layer_b_host, const float* freq, const float* phase,
                   const float* final_w, const float* final_b, int sigmoid_rgb, int precision, const float* u_jitter,
                   const float* noise_coarse, const float* u_resample, const float* noise_final, int hierarchical, float noise_std,
                   int clamp_mode, int white_back, int last_back, void* workspace, size_t workspace_bytes, float* pixels,
                   float* depth, cng_stream_t stream) {
  CNG_REQUIRE(B >= 0 && img_w >= 1 && img_h >= 1 && S >= (hierarchical ? 3 : 2), CNG_ERR_INVALID_ARGUMENT,
              "render_fwd: B=%d img=%dx%d S=%d", B, img_w, img_h, S);
  if (B == 0) return CNG_OK;
  CNG_REQUIRE(pixels && depth && workspace, CNG_ERR_INVALID_ARGUMENT, "render_fwd: NULL output / workspace");
  CNG_REQUIRE(!hierarchical || u_resample, CNG_ERR_INVALID_ARGUMENT, "render_fwd: hierarchical sampling needs u_resample");
  CNG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, CNG_ERR_INVALID_ARGUMENT, "render_fwd: workspace not 256-byte aligned");
  const long long R = static_cast<long long>(img_w) * img_h;
  const cng::RenderLayout lay = cng::render_layout(B, R, S, C, HID, L, hierarchical, precision);
  CNG_REQUIRE(workspace_bytes >= lay.total, CNG_ERR_WORKSPACE, "render_fwd: workspace %zu < %zu bytes", workspace_bytes, lay.total);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* feat = reinterpret_cast<float*>(ws + lay.feat);
  float* t_c = reinterpret_cast<float*>(ws + lay.t_c);
  float* rs_c = reinterpret_cast<float*>(ws + lay.rs_c);
  void* mlp_ws = ws + lay.mlp;
  const size_t mlp_bytes = cng_film_siren_workspace_bytes(B, C, HID, L, precision);
  const long long N = R * S;

  const bool fused = cng::fused_gather(C, HID, precision);
  float* pts = feat;                                   // fused mode: the feat slot of the workspace holds the positions [B, N, 3]
  if (fused) {
    if (int e = cng_raymarch_gather_coarse(vol_ndhwc, vol_item_stride, B, C, D, H, W, cam2world, rays_d_cam, t_lin, u_jitter, img_w, img_h, S,
                                           nullptr, t_c, pts, stream)) return e;
    if (int e = cng_film_siren_fwd_gather(vol_ndhwc, vol_item_stride, D, H, W, pts, B, N, C, HID, L, layer_w_host, layer_b_host, freq, phase,
                                          final_w, final_b, sigmoid_rgb, precision, mlp_ws, mlp_bytes, rs_c, stream)) return e;
  } else {
    if (int e = cng_raymarch_gather_coarse(vol_ndhwc, vol_item_stride, B, C, D, H, W, cam2world, rays_d_cam, t_lin, u_jitter, img_w, img_h, S,
                                           feat, t_c, nullptr, stream)) return e;
    if (int e = cng_film_siren_fwd(feat, B, N, C, HID, L, layer_w_host, layer_b_host, freq, phase, final_w, final_b, sigmoid_rgb, precision,
                                   mlp_ws, mlp_bytes, rs_c, stream)) return e;
  }
  if (!hierarchical)
    return cng_merge_composite(nullptr, rs_c, nullptr, t_c, noise_final ? noise_final : noise_coarse, rays_d_cam, B, static_cast<int>(R), S,
                               noise_std, clamp_mode, white_back, last_back, pixels, depth, nullptr, nullptr, nullptr, stream);
  float* w_c = reinterpret_cast<float*>(ws + lay.w_c);
  float* t_f = reinterpret_cast<float*>(ws + lay.t_f);
  float* rs_f = reinterpret_cast<float*>(ws + lay.rs_f);
  // coarse weights for the resampling: no white / last background (generators.py:115-121)
  if (int e = cng_composite_fwd(rs_c, t_c, noise_coarse, static_cast<long long>(B) * R, S, noise_std, clamp_mode, 0, 0, nullptr, nullptr, w_c,
                                stream)) return e;
  if (int e = cng_resample_from_coarse(t_c, w_c, u_resample, static_cast<long long>(B) * R, S, t_f, nullptr, stream)) return e;
  if (fused) {
    if (int e = cng_raymarch_gather_fine(vol_ndhwc, vol_item_stride, B, C, D, H, W, cam2world, rays_d_cam, t_f, img_w, img_h, S, nullptr, pts,
                                         stream)) return e;
    if (int e = cng_film_siren_fwd_gather(vol_ndhwc, vol_item_stride, D, H, W, pts, B, N, C, HID, L, layer_w_host, layer_b_host, freq, phase,
                                          final_w, final_b, sigmoid_rgb, precision, mlp_ws, mlp_bytes, rs_f, stream)) return e;
  } else {
    if (int e = cng_raymarch_gather_fine(vol_ndhwc, vol_item_stride, B, C, D, H, W, cam2world, rays_d_cam, t_f, img_w, img_h, S, feat, nullptr,
                                         stream)) return e;
    if (int e = cng_film_siren_fwd(feat, B, N, C, HID, L, layer_w_host, layer_b_host, freq, phase, final_w, final_b, sigmoid_rgb, precision,
                                   mlp_ws, mlp_bytes, rs_f, stream)) return e;
  }
  return cng_merge_composite(rs_f, rs_c, t_f, t_c, noise_final, rays_d_cam, B, static_cast<int>(R), S, noise_std, clamp_mode, white_back,
                             last_back, pixels, depth, nullptr, nullptr, nullptr, stream);
}

}  // extern "C"
