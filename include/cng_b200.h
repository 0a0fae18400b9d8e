/*
 * cng_b200.h -- C ABI of the B200-native rendering hot path of the conditioned pi-GAN style
 * NeRF-GAN (drop-in for zzhuolun/conditioned-nerf-gan's ImplicitGenerator3d.forward).
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; the
 * boundary it exposes is the Python class contract `ImplicitGenerator3d.forward(z, cam2worlds,
 * **curriculum)`.  Each entry point below replaces one block of stock-ATen launches on that
 * path and cites the reference lines it replaces (paths relative to the reference checkout).
 * The Python host mirror (conditioned_nerf_gan_b200/generators/*.py) binds these with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name ends in `_host`;
 *   - tensors are dense, row-major, fp32 unless stated; shapes are written [outer, ..., inner];
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions return 0 on success, a negative CNG_ERR_* code for argument errors, or a
 *     positive cudaError_t; cng_last_error() gives a thread-local message.  Nothing is thrown,
 *     nothing falls back to the CPU: without a CUDA device every compute call fails with
 *     CNG_ERR_NO_DEVICE (or the CUDA error) and leaves the outputs untouched;
 *   - no call synchronises the device; outputs are valid in stream order.
 *
 * Symbols: R = rays per image (img_h * img_w), S = coarse samples per ray (`num_steps`),
 * C = feature channels, (D, H, W) = feature-volume extent (z, y, x), L = FiLM layers,
 * HID = hidden width (256).
 */
#ifndef CNG_B200_H_
#define CNG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNG_ABI_VERSION 1

#if defined(__GNUC__)
#define CNG_API __attribute__((visibility("default")))
#else
#define CNG_API
#endif

typedef void* cng_stream_t;

enum {
  CNG_OK = 0,
  CNG_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, non-positive size, misaligned buffer        */
  CNG_ERR_UNSUPPORTED = -2,      /* shape outside what the kernels are built for              */
  CNG_ERR_NO_DEVICE = -3,        /* no sm_100 device visible                                  */
  CNG_ERR_WORKSPACE = -4         /* workspace too small (see *_workspace_bytes)               */
};

/* reference: `clamp_mode` kwarg of fancy_integration, volumetric_rendering.py:41-46 */
enum { CNG_CLAMP_RELU = 0, CNG_CLAMP_SOFTPLUS = 1 };

/* arithmetic of the FiLM-SIREN contractions */
enum {
  CNG_PREC_FP32 = 0,   /* fp32 FFMA tiles (exact-mode; reference inference is pure fp32)       */
  CNG_PREC_BF16 = 1,   /* tcgen05 bf16 x bf16 -> fp32 TMEM accumulators; layer 0 split-bf16    */
  CNG_PREC_FP16 = 2    /* same kernel and rate with fp16 operands (the reference's autocast
                          dtype): 11-bit instead of 8-bit significands, ~8x smaller error     */
};

CNG_API int cng_abi_version(void);
CNG_API const char* cng_last_error(void);
/* 0 if an sm_100 device is current and the kernels can launch, else an error code. */
CNG_API int cng_device_check(void);

/* ------------------------------------------------------------------------------------------
 * a1 helper (host): the per-pixel camera-space directions and the coarse distances.
 * Replaces get_initial_rays_trig, generators/volumetric_rendering.py:73-100 (host part).
 * rays_d_cam_host [img_h*img_w, 3], t_lin_host [S].  Ray p = row*img_w + col, x = lin(-1,1,W)[col],
 * y = lin(-1,1,H)[row], z = 1/tan(fov/2) (float64 tan), normalised.  The Python mirror instead
 * builds these tables with the same torch ops as the reference so they agree to the last bit.
 * ---------------------------------------------------------------------------------------- */
CNG_API int cng_camera_tables_host(int img_w, int img_h, int S, double fov_deg, double ray_start,
                           double ray_end, float* rays_d_cam_host, float* t_lin_host);

/* ------------------------------------------------------------------------------------------
 * a4 prologue: NCDHW -> NDHWC so that one trilinear corner is one contiguous C*4-byte line.
 * The reference hands F.grid_sample an NCDHW volume (generators/siren.py:555-571).
 * ---------------------------------------------------------------------------------------- */
CNG_API int cng_volume_to_channels_last(const float* vol_ncdhw, float* vol_ndhwc, int B, int C, int D,
                                int H, int W, cng_stream_t stream);
/* The same for a volume held in fp16 (the dtype the encoder emits under the trainer's autocast, utils.py:643-647; half the
 * host->device bytes when volumes are streamed from host memory): widened to fp32 in the pass that re-lays it. */
CNG_API int cng_volume_f16_to_channels_last(const void* vol_ncdhw_f16, float* vol_ndhwc, int B, int C, int D,
                                    int H, int W, cng_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K1 coarse: sample generation + stratified jitter + cam->world + trilinear gather, fused.
 * Replaces get_initial_rays_trig (volumetric_rendering.py:73-100), perturb_points (:103-110),
 * transform_sampled_points (:113-199) and the F.grid_sample + reshape/permute block of every
 * feature-volume SIREN (generators/siren.py:555-571).
 *   vol_ndhwc  [B, D, H, W, C]      cam2world [B, 4, 4]
 *   vol_item_stride = C*D*H*W floats, or 0 when all B cameras look at ONE object whose volume is
 *                     vol_ndhwc[0] (video rendering, inference.py:441-486)
 *   rays_d_cam [R, 3]               t_lin [S]
 *   u_jitter   [B, R, S] uniform draws (the reference's torch.rand at :106), NULL = no jitter
 *   feat       [B, R, S, C] out     t_out [B, R, S] out (jittered distances)
 *   points_out [B, R, S, 3] out, may be NULL (world-space sample positions, for taps/tests)
 * Requires C % 4 == 0, C <= 128, S >= 2, img_w * img_h == R.
 * ---------------------------------------------------------------------------------------- */
CNG_API int cng_raymarch_gather_coarse(const float* vol_ndhwc, long long vol_item_stride, int B, int C, int D, int H, int W,
                               const float* cam2world, const float* rays_d_cam,
                               const float* t_lin, const float* u_jitter, int img_w, int img_h,
                               int S, float* feat, float* t_out, float* points_out,
                               cng_stream_t stream);

/* K1 fine: p = o + d * t_fine in world space (generators/generators.py:138-145) + gather.
 *   t_fine [B, R, S];  feat [B, R, S, C] out;  points_out [B, R, S, 3] out or NULL. */
CNG_API int cng_raymarch_gather_fine(const float* vol_ndhwc, long long vol_item_stride, int B, int C, int D, int H, int W,
                             const float* cam2world, const float* rays_d_cam,
                             const float* t_fine, int img_w, int img_h, int S, float* feat,
                             float* points_out, cng_stream_t stream);

/* Gather at caller-supplied world points: the lookup inside `siren(points, z, img_size,
 * num_steps)` (generators/siren.py:555-571) when it is called directly (extract_shapes.py:63-68).
 *   points [B, N, 3];  feat [B, N, C] out;  corner_idx [B, N, 3] int32 out or NULL: floor of the
 *   clamped continuous voxel index per axis (x, y, z) -- exposed so tests can pin the indexing. */
CNG_API int cng_gather_points(const float* vol_ndhwc, int B, int C, int D, int H, int W,
                      const float* points, long long N, float* feat, int32_t* corner_idx,
                      cng_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2: FiLM-SIREN MLP, all layers fused, activations never leave the SM.
 * Replaces, per layer, addmm + expand + mul + add + sin (FiLMLayer.forward, siren.py:146-160;
 * loops at :573-577, :661-665, :820-824, :1058-1062) and the head nn.Linear(256,4) +
 * _sigmoid_rgb (:579, :1227-1234).
 *   feat [B, N, C]; layer_w_host / layer_b_host: HOST arrays of L device pointers to
 *   nn.Linear weight [HID, K_l] (K_0 = C, else HID) and bias [HID];
 *   freq, phase [B, L*HID] (freq already `*15+30`, siren.py:550-553);
 *   final_w [4, HID], final_b [4]; rgb_sigma [B, N, 4] out.
 * CNG_PREC_BF16 / CNG_PREC_FP16 need HID == 256, C == 32 and a workspace of
 * cng_film_siren_workspace_bytes().
 * ---------------------------------------------------------------------------------------- */
/* a5: FiLM parameters of the mapping network (nn.Linear(z_dim, 2*L*HID) on the global feature, siren.py:550-553):
 *   out = map_w [n_out, z_dim] . global_feature [B, z_dim] + map_b;  freq [B, n_out/2] = out[:, :n_out/2] * 15 + 30,
 *   phase [B, n_out/2] = out[:, n_out/2:].  Fixed accumulation order per output: independent of the batch size. */
CNG_API int cng_film_parameters(const float* global_feature, const float* map_w, const float* map_b, int B,
                        int z_dim, int n_out, float* freq, float* phase, cng_stream_t stream);

CNG_API size_t cng_film_siren_workspace_bytes(int B, int C, int HID, int L, int precision);
CNG_API int cng_film_siren_fwd(const float* feat, int B, long long N, int C, int HID, int L,
                       const float* const* layer_w_host, const float* const* layer_b_host,
                       const float* freq, const float* phase, const float* final_w,
                       const float* final_b, int sigmoid_rgb, int precision, void* workspace,
                       size_t workspace_bytes, float* rgb_sigma, cng_stream_t stream);

/* K2 with residual blocks (TALLSIREN_dRes, generators/siren.py:218-230, 333-408: y = sin(x + fc2(sin(fc1 x)))).  The
 * network is passed as its flat list of L linear layers; bit l of res_save_mask = layer l's output is kept as "x",
 * bit l of res_add_mask = the kept x is added to layer l's pre-activation (a layer may do both: add, then keep).
 * res_scratch: cng_film_siren_res_scratch_bytes() of device memory for the 16-bit modes (the kept activations of the
 * tiles in flight, fp32, L2-resident), unused for CNG_PREC_FP32.  Masks 0 = cng_film_siren_fwd. */
/* K2 with the trilinear lookup (a4) fused into its prologue: every 128-point tile looks its points [B, N, 3] (world space, as
 * written by cng_raymarch_gather_{coarse,fine} with feat == NULL) up in the NDHWC volume itself, so the gathered features
 * [B, N, 32] never go to HBM (403 MB written + read per pass at BASELINE configs[1]).  Tensor-core precisions, plain FiLM
 * networks, C == 32.  Same arithmetic as cng_gather_points followed by cng_film_siren_fwd: bit-identical output. */
CNG_API int cng_film_siren_fwd_gather(const float* vol_ndhwc, long long vol_item_stride, int D, int H, int W,
                              const float* points, int B, long long N, int C, int HID, int L,
                              const float* const* layer_w_host, const float* const* layer_b_host,
                              const float* freq, const float* phase, const float* final_w,
                              const float* final_b, int sigmoid_rgb, int precision, void* workspace,
                              size_t workspace_bytes, float* rgb_sigma, cng_stream_t stream);
CNG_API size_t cng_film_siren_res_scratch_bytes(void);
CNG_API int cng_film_siren_fwd_res(const float* feat, int B, long long N, int C, int HID, int L,
                           const float* const* layer_w_host, const float* const* layer_b_host,
                           const float* freq, const float* phase, const float* final_w, const float* final_b,
                           int sigmoid_rgb, int precision, unsigned res_save_mask, unsigned res_add_mask,
                           void* workspace, size_t workspace_bytes, void* res_scratch, size_t res_scratch_bytes,
                           float* rgb_sigma, cng_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3: alpha compositing, one warp per ray, exclusive-cumprod transmittance by warp shuffles.
 * Replaces fancy_integration (volumetric_rendering.py:18-70).
 *   rgb_sigma [n_rays, S, 4]; t [n_rays, S]; noise [n_rays, S] (the reference's randn at :39)
 *   or NULL when noise_std == 0; outputs rgb [n_rays, 3], dist [n_rays], weights [n_rays, S]
 *   (any may be NULL).  S <= 1024.
 * ---------------------------------------------------------------------------------------- */
CNG_API int cng_composite_fwd(const float* rgb_sigma, const float* t, const float* noise,
                      long long n_rays, int S, float noise_std, int clamp_mode, int white_back,
                      int last_back, float* rgb, float* dist, float* weights,
                      cng_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4: inverse-CDF importance resampling, CDF + binary search in shared memory.
 * Replaces sample_pdf (volumetric_rendering.py:297-342, det=False).
 *   bins [n, M+1]; weights [n, M]; u [n, K] uniform draws (torch.rand at :321);
 *   samples [n, K] out; inds [n, K] int64 out or NULL (= torch.searchsorted(cdf, u), :324).
 * Bit-exact against the oracle: sum and CDF accumulate sequentially in float64 and round to
 * fp32 per element (what torch.cumsum does on CPU); the search and the lerp use IEEE fp32 ops
 * in the reference's order.  M <= 2047.
 * ---------------------------------------------------------------------------------------- */
CNG_API int cng_sample_pdf(const float* bins, const float* weights, const float* u, long long n, int M,
                   int K, float eps, float* samples, int64_t* inds, cng_stream_t stream);

/* The sample_pdf call site fused (generators/generators.py:123-136): bins = midpoints of
 * t_coarse, weights = (w[1:-1] + 1e-5), N_importance = S.
 *   t_coarse, weights [n, S]; u [n, S]; t_fine [n, S] out; inds [n, S] int64 out or NULL. */
CNG_API int cng_resample_from_coarse(const float* t_coarse, const float* weights, const float* u,
                             long long n, int S, float* t_fine, int64_t* inds,
                             cng_stream_t stream);

/* a11 alone: the order of torch.sort over cat([fine, coarse]) along the sample axis (generators/generators.py:163-167),
 * stable with the fine samples first among equal distances.
 *   t_fine, t_coarse [n_rays, S]; order [n_rays, 2S] int32 out (index into the fine-first concatenation) and / or
 *   t_sorted [n_rays, 2S] out; either may be NULL.   2S <= 512. */
CNG_API int cng_merge_sort(const float* t_fine, const float* t_coarse, long long n_rays, int S, int32_t* order,
                   float* t_sorted, cng_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3 final: merge coarse+fine by depth (stable: fine first), composite, format the image.
 * Replaces cat/sort/gather (generators/generators.py:163-167), the final fancy_integration
 * (:172-180), the NCHW permute `*2-1` (:182-183) and distance2depth (:185-186,
 * volumetric_rendering.py:345-356).
 *   rgb_sigma_fine / _coarse [B*R, S, 4]; t_fine / t_coarse [B*R, S]; noise [B*R, 2S] or NULL;
 *   rays_d_cam [R, 3]; outputs pixels [B, 3, img_h, img_w], depth [B, img_h, img_w];
 *   optional taps rgb [B*R, 3], dist [B*R], order [B*R, 2S] int32 (index into the fine-first
 *   concatenation), may be NULL.   2S <= 512.
 * With rgb_sigma_fine == NULL the coarse samples alone are composited (hierarchical_sample
 * False) and noise is [B*R, S].
 * ---------------------------------------------------------------------------------------- */
CNG_API int cng_merge_composite(const float* rgb_sigma_fine, const float* rgb_sigma_coarse,
                        const float* t_fine, const float* t_coarse, const float* noise,
                        const float* rays_d_cam, int B, int R, int S, float noise_std,
                        int clamp_mode, int white_back, int last_back, float* pixels,
                        float* depth, float* rgb, float* dist, int32_t* order,
                        cng_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Backward: what autograd computes for generators/generators.py:102-180
 * when the trainer calls loss.backward() (utils.py:711).  Sample positions carry no gradient
 * (built under torch.no_grad(), generators.py:57,111).
 * ---------------------------------------------------------------------------------------- */

/* Backward of cng_merge_composite: d_pixels [B,3,img_h,img_w] and/or d_depth [B,img_h,img_w]
 * (either may be NULL) -> d_rgb_sigma_fine / d_rgb_sigma_coarse [B*R, S, 4] (every element is
 * written).  Forward inputs are passed again; the merge order, alpha and transmittance are
 * recomputed.  With rgb_sigma_fine == NULL it is the backward of a plain fancy_integration
 * (volumetric_rendering.py:18-70) over the coarse samples. */
CNG_API int cng_merge_composite_bwd(const float* rgb_sigma_fine, const float* rgb_sigma_coarse,
                            const float* t_fine, const float* t_coarse, const float* noise,
                            const float* rays_d_cam, const float* d_pixels, const float* d_depth,
                            int B, int R, int S, float noise_std, int clamp_mode, int white_back,
                            int last_back, float* d_rgb_sigma_fine, float* d_rgb_sigma_coarse,
                            cng_stream_t stream);

/* Backward of cng_composite_fwd (plain fancy_integration, volumetric_rendering.py:18-70): d_rgb [n_rays, 3] and / or
 * d_dist [n_rays] (either may be NULL; the `weights` output carries no gradient here) -> d_rgb_sigma [n_rays, S, 4]. */
CNG_API int cng_composite_bwd(const float* rgb_sigma, const float* t, const float* noise, const float* d_rgb,
                      const float* d_dist, long long n_rays, int S, float noise_std, int clamp_mode,
                      int white_back, int last_back, float* d_rgb_sigma, cng_stream_t stream);

/* Backward of the trilinear lookup w.r.t. the volume (F.grid_sample backward, siren.py:555-571):
 * dvol_ndhwc [B,D,H,W,C] += trilinear weights * dfeat [B,N,C] at points [B,N,3] (vector
 * red.global.add).  The caller zero-fills dvol_ndhwc first. */
CNG_API int cng_scatter_points(float* dvol_ndhwc, int B, int C, int D, int H, int W, const float* points,
                       long long N, const float* dfeat, cng_stream_t stream);

/* NDHWC -> NCDHW: the volume gradient back in the layout of the encoder's output. */
CNG_API int cng_volume_from_channels_last(const float* vol_ndhwc, float* vol_ncdhw, int B, int C, int D,
                                  int H, int W, cng_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * The whole forward of ImplicitGenerator3d.forward (generators/generators.py:33-187) in one call: K1 coarse, K2, K3,
 * K4, K1 fine, K2, K3' sequenced on `stream` with every intermediate in `workspace`
 * (cng_render_workspace_bytes, 256-byte aligned).  Inputs: the NDHWC volume (vol_item_stride as in
 * cng_raymarch_gather_coarse), the camera tables, the SIREN parameters with freq / phase from cng_film_parameters,
 * and the four random draws of the reference in its order: u_jitter [B,R,S], noise_coarse [B,R,S] (or NULL),
 * u_resample [B*R,S], noise_final [B,R,2S] (or NULL; [B,R,S] when hierarchical == 0).
 * Outputs: pixels [B,3,img_h,img_w], depth [B,img_h,img_w].
 * ---------------------------------------------------------------------------------------- */
CNG_API size_t cng_render_workspace_bytes(int B, int img_w, int img_h, int S, int C, int HID, int L,
                                  int hierarchical, int precision);
CNG_API int cng_render_fwd(const float* vol_ndhwc, long long vol_item_stride, int B, int C, int D, int H, int W,
                   const float* cam2world, const float* rays_d_cam, const float* t_lin, int img_w,
                   int img_h, int S, int HID, int L, const float* const* layer_w_host,
                   const float* const* layer_b_host, const float* freq, const float* phase,
                   const float* final_w, const float* final_b, int sigmoid_rgb, int precision,
                   const float* u_jitter, const float* noise_coarse, const float* u_resample,
                   const float* noise_final, int hierarchical, float noise_std, int clamp_mode,
                   int white_back, int last_back, void* workspace, size_t workspace_bytes,
                   float* pixels, float* depth, cng_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * a14: backward of the FiLM-SIREN MLP (autograd through FiLMLayer.forward x L + head, generators/siren.py:146-160,
 * 573-579; what loss.backward() runs at utils.py:711), all contractions on tcgen05 / TMEM (film_siren_bwd_tc.cu).
 * T = ceil(points / 128) tiles.  Dump formats (device buffers the caller owns):
 *   x_dump    [L][T][65536]  output of layer l, x_{l+1} = sin(u_l), as 128-point operand tile images
 *                            ([4 K-blocks of 64 columns][128 rows][128 B], 128-byte swizzle: element (row, k) of a block at
 *                            row*128 + (((k>>3) ^ (row&7)) << 4) + (k&7)*2), in the operand format of `precision`
 *   g_dump    [L][T][G]      g_l = cos(u_l) in the epilogue's register order (row = 32 q + lane), G = 128 * 256 * bits / 8 bytes with
 *                            bits = cng_film_siren_g_dump_bits():
 *                              16 (default): fp16, [32-column block cc 8][row quarter q 4][piece i 4][lane 32] x 16 B, columns
 *                                  32 cc + 8 i .. + 7 (G = 65536);
 *                               8 (CNG_G_DUMP_BITS=8): codes round(127 g) + 128, [cc 8][q 4][half h 2][lane 32] x 16 B, byte e = column
 *                                  32 cc + 16 h + e (G = 32768).  Halves the dump's traffic (the training forward is bound by its
 *                                  HBM writes: MLP backward 4.9 -> 4.5 ms per 1 M points) for a quantisation error <= 1/254 per
 *                                  element: gradients of the 16 x 16 test scenes move from 0.9 % to 1.2 % relative L2.
 *                            The FiLM frequency is not applied elementwise:
 *                            with dz'_l = dy_l * cos(u_l) the chain is dy_{l-1} = dz'_l (diag(freq_l) W_l) and
 *                            dW_l = diag(freq_l) dz'_l^T x_l, db_l = freq_l * colsum(dz'_l), dphase_l = colsum(dz'_l),
 *                            dfreq_l = rowsum(W_l * dz'_l^T x_l) + b_l * colsum(dz'_l)
 *   feat_dump [T][16384]     the layer-0 operand block [x_hi(32) | x_lo(32)] (same swizzle)
 *   dz_dump   [L][T][65536]  dz'_l = dy_l * g_l as bf16 tile images (written by the dgrad chain, read by the weight gradient)
 * ---------------------------------------------------------------------------------------- */
/* Training-mode forward of K2 (the backward's activation recompute): the fused tcgen05 kernel of cng_film_siren_fwd(_res) that
 * ALSO writes the three dumps above (one bulk store per tile-layer for x, 16-byte register stores in the epilogue's order for g).  `precision`:
 * CNG_PREC_BF16 or CNG_PREC_FP16 (operand format of the recompute and of x_dump).  With B > 1 the tile index runs over items:
 * T = B * ceil(N / 128).  Residual masks / scratch as in cng_film_siren_fwd_res (0 / NULL for plain networks); g_l is then the
 * derivative at the pre-activation INCLUDING the re-added block input. */
CNG_API int cng_film_siren_fwd_train(const float* feat, int B, long long N, int C, int HID, int L,
                             const float* const* layer_w_host, const float* const* layer_b_host,
                             const float* freq, const float* phase, const float* final_w,
                             const float* final_b, int sigmoid_rgb, int precision, unsigned res_save_mask,
                             unsigned res_add_mask, void* workspace, size_t workspace_bytes, void* res_scratch,
                             size_t res_scratch_bytes, float* rgb_sigma, void* x_dump, void* g_dump,
                             void* feat_dump, cng_stream_t stream);
/* Operand images (bf16) of (diag(freq_l) W_l)^T and of the head for the dgrad chain.  freq [L*HID] of the item, or NULL for
 * plain W_l.  `images`: cng_film_siren_wt_image_bytes(L) bytes, 16-byte aligned. */
CNG_API size_t cng_film_siren_wt_image_bytes(int L);
/* Bits per element of g_dump (16 or 8, see "Dump formats"): process-wide, read by cng_film_siren_fwd_train, cng_film_siren_dgrad
 * and cng_film_siren_bwd at call time; a dump must be consumed in the format it was written in. */
CNG_API int cng_film_siren_g_dump_bits(void);
CNG_API int cng_film_siren_wt_images(const float* const* layer_w_host, const float* freq, const float* final_w, int C,
                             int HID, int L, void* images, cng_stream_t stream);
/* The dgrad chain of P points, all layers fused per 128-point tile (gradients stay in TMEM / shared memory between layers):
 *   d_o = d_out (* rgb (1 - rgb) when sigmoid_rgb, with `out` the forward output), d_final_b_acc [4] += colsum(d_o),
 *   dy = d_o Wf, then for l = L-1 .. 0: dz_l = dy * g_l (-> dz_dump), dy = dz_l W_l; d_feat [P, 32] = dz_0 W_0 (written).
 * A kept activation of a residual block additionally receives the adding layer's dz (masks / scratch as in
 * cng_film_siren_fwd_res; only patterns where an add sits exactly two layers after its kept layer: CNG_ERR_UNSUPPORTED
 * otherwise).  g_layer_stride_tiles (here) / x_layer_stride_tiles (below): tiles per layer of a dump that holds more tiles than
 * this call processes -- the dumps of a whole batch written by ONE cng_film_siren_fwd_train call and consumed item by item, g_dump /
 * x_dump pointing at the item's first tile; 0 = this call's own tile count. */
CNG_API int cng_film_siren_dgrad(const float* d_out, const float* out, int sigmoid_rgb, long long P, int L,
                         const void* wt_images, const void* g_dump, long long g_layer_stride_tiles, void* dz_dump,
                         float* d_feat, float* d_final_b_acc, unsigned res_save_mask, unsigned res_add_mask,
                         void* res_scratch, size_t res_scratch_bytes, cng_stream_t stream);
/* Weight gradients as a split-K contraction over the points: d_w_acc_host[l] [HID, K_l] += dz_l^T x_l (x_0 = the features,
 * hi + lo), colsum_acc [L, HID] += column sums of dz_l.  Both operands are read from the tile images with MN-major
 * descriptors; fp32 accumulation in TMEM, one red.global.add flush per CTA and layer.  x_is_fp16: format of x_dump / feat_dump. */
CNG_API int cng_film_siren_wgrad(const void* dz_dump, const void* x_dump, long long x_layer_stride_tiles,
                         const void* feat_dump, long long P, int L, int x_is_fp16, float* const* d_w_acc_host,
                         float* colsum_acc, cng_stream_t stream);
/* Head weights: d_final_w_acc [4, HID] += d_o^T x_L, d_o as in cng_film_siren_dgrad, x_L = the tile images of the last layer's
 * output (x_dump + (L - 1) * stride * 65536 [+ the item's tile offset]). */
CNG_API int cng_film_siren_head_wgrad(const float* d_out, const float* out, int sigmoid_rgb, const void* x_last_tiles,
                              long long P, int x_is_fp16, float* d_final_w_acc, cng_stream_t stream);
/* The whole MLP backward of one chunk of points of ONE batch item in one call: recompute with dumps (fp16 operands), W^T
 * images, dgrad chain, weight gradients, head weights.
 *   feat [P, C]; d_out [P, 4] gradient w.r.t. rgb_sigma;
 *   layer_w_host / layer_b_host: HOST arrays of L device pointers (fp32); freq, phase [L*HID] of this item; final_w [4, HID], final_b [4];
 *   outputs: d_feat [P, C] (written); accumulated (+=): d_w_acc_host[l] [HID, K_l] fp32 = dz'_l^T x_l and colsum_acc [L, HID] =
 *   column sums of dz'_l, both WITHOUT the FiLM frequency (see g_dump above: the host forms dW = freq * dW', d_bias =
 *   freq * colsum', d_phase = colsum', d_freq = rowsum(W * dW') + b * colsum'), d_final_w_acc [4, HID], d_final_b_acc [4].
 *   workspace: cng_film_siren_bwd_workspace_bytes(P, C, HID, L), 256-byte aligned.  Residual masks / scratch as in
 *   cng_film_siren_fwd_res.  HID == 256, C == 32, P < 2^31. */
CNG_API size_t cng_film_siren_bwd_workspace_bytes(long long P, int C, int HID, int L);
CNG_API int cng_film_siren_bwd(const float* feat, const float* d_out, long long P, int C, int HID, int L,
                       const float* const* layer_w_host, const float* const* layer_b_host,
                       const float* freq, const float* phase, const float* final_w, const float* final_b,
                       int sigmoid_rgb, unsigned res_save_mask, unsigned res_add_mask, void* workspace,
                       size_t workspace_bytes, void* res_scratch, size_t res_scratch_bytes, float* d_feat,
                       float* const* d_w_acc_host, float* colsum_acc, float* d_final_w_acc, float* d_final_b_acc,
                       cng_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * GroupNorm of the 3D U-Net encoder (generators/unet3d.py:21-132, nn.GroupNorm in every SingleConv) on
 * channels-last volumes x [N, S, C] (S = D*H*W; the memory of a channels_last_3d [N,C,D,H,W] tensor), the layout in which
 * the encoder emits the feature volume the gather kernel reads.  dtype: 0 fp32, 1 fp16, 2 bf16 (x, y, dy, dx share it;
 * statistics in fp32 / fp64).  gamma / beta [C] fp32 or NULL.  C <= 1024, G <= 64.
 *   fwd: y = (x - mean) * rstd * gamma + beta; mean, rstd [N, G] are written (the backward reads them);
 *        sums_scratch: N*G*2 doubles.
 *   bwd: dx from dy; ds_scratch / db_scratch [N, C] fp32 receive sum_s dy * xhat and sum_s dy per item and channel
 *        (d_gamma = their sum over N of ds, d_beta = of db: left to the caller).
 * ---------------------------------------------------------------------------------------- */
CNG_API int cng_group_norm_fwd(const void* x, int dtype, int N, long long S, int C, int G, const float* gamma,
                       const float* beta, float eps, void* y, float* mean, float* rstd, double* sums_scratch,
                       cng_stream_t stream);
CNG_API int cng_group_norm_bwd(const void* dy, const void* x, int dtype, int N, long long S, int C, int G,
                       const float* gamma, const float* mean, const float* rstd, void* dx, float* ds_scratch,
                       float* db_scratch, cng_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CNG_B200_H_ */
