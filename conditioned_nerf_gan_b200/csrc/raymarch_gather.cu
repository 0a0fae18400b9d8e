// K1: fused ray-march (sample generation, stratified jitter, cam->world) + trilinear gather.
//
// Replaces get_initial_rays_trig / perturb_points / transform_sampled_points
// (generators/volumetric_rendering.py:73-199), the fine-point synthesis
// (generators/generators.py:138-145) and F.grid_sample(bilinear, border, align_corners=False) +
// reshape/permute of every feature-volume SIREN (generators/siren.py:555-571).
//
// Data layout: the volume is channels-last [B, D, H, W, C] (cng_volume_to_channels_last), so one
// trilinear corner of a C=32 volume is a single 128-byte line.  A sample point is served by a
// group of 8 lanes, each loading one float4 (4 channels) per corner: a warp instruction reads four
// full lines, and the 8 corner loads of a thread are independent (8 x 16 B in flight per lane).
// A block walks an 8x4 pixel patch along the ray, so the 32 points of one iteration are spatial
// neighbours (0.1-0.75 voxels apart) and re-use each other's lines through L1; the compulsory
// HBM traffic is one read of the volume (which fits L2: 33.5 MB per 64^3 x 32 object).
// No positions, no homogeneous temporaries and no [B,C,N] gather result ever go to HBM: only
// feat[B,R,S,C] (+ the jittered distances).
//
// Index arithmetic follows ATen's grid_sampler_3d (border padding, align_corners=False) with
// explicit round-to-nearest intrinsics so no FMA contraction changes a floor():
//   g = p / 0.6;  i = ((g + 1) * size - 1) / 2;  i = min(size-1, max(i, 0));  i0 = floor(i);
//   w_lo = (i0 + 1) - i, w_hi = i - i0;  corners accumulated in ATen's order tnw..bse.
#include <cuda_fp16.h>

#include "cng_common.cuh"
#include "trilinear.cuh"

#ifndef CNG_K1_UNROLL
#define CNG_K1_UNROLL 2         // rounds of phase B unrolled together (A/B knob)
#endif
#ifndef CNG_K1_MIN_BLOCKS_C8
#define CNG_K1_MIN_BLOCKS_C8 4  // ... of the 32-channel instantiation: it fits 64 registers without spilling (measured: c2 step 6.44 -> 6.41 ms)
#endif
#ifndef CNG_K1_MIN_BLOCKS
#define CNG_K1_MIN_BLOCKS 3     // resident blocks per SM the register allocation aims at (A/B knob)
#endif

namespace cng {

constexpr int kK1Unroll = CNG_K1_UNROLL;
constexpr int kLanesPerPoint = 8;
constexpr int kPointsPerBlock = 32;      // 8x4 pixel patch
constexpr int kTileW = 8, kTileH = 4;

struct Corner {
  int x0, y0, z0;
  float fx1, fx0, fy1, fy0, fz1, fz0;    // weights of the high / low corner per axis
};

// Accumulate the 8 corners for channel group `cg4` (float4 index) of one point.
__device__ __forceinline__ float4 trilinear_c4(const float4* __restrict__ vol, int D, int H, int W, int C4,
                                               float px, float py, float pz, int cg4, int* idx_out) {
  int x0, y0, z0;
  float xl, xh, yl, yh, zl, zh;
  axis_index(px, W, x0, xl, xh);
  axis_index(py, H, y0, yl, yh);
  axis_index(pz, D, z0, zl, zh);
  if (idx_out) { idx_out[0] = x0; idx_out[1] = y0; idx_out[2] = z0; }
  const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1), z1 = min(z0 + 1, D - 1);
  // out-of-range high corners carry weight exactly 0 in ATen (they are skipped); with border
  // clamping i <= size-1, so i0 == size-1 implies w_hi == 0 and the clamped load is harmless.
  const bool xin = x0 + 1 <= W - 1, yin = y0 + 1 <= H - 1, zin = z0 + 1 <= D - 1;
  const size_t sx = C4, sy = static_cast<size_t>(W) * C4, sz = static_cast<size_t>(H) * W * C4;
  const float4* b = vol + cg4;
  // issue all 8 loads before any use
  const float4 v000 = __ldg(b + z0 * sz + y0 * sy + x0 * sx);
  const float4 v001 = __ldg(b + z0 * sz + y0 * sy + x1 * sx);
  const float4 v010 = __ldg(b + z0 * sz + y1 * sy + x0 * sx);
  const float4 v011 = __ldg(b + z0 * sz + y1 * sy + x1 * sx);
  const float4 v100 = __ldg(b + z1 * sz + y0 * sy + x0 * sx);
  const float4 v101 = __ldg(b + z1 * sz + y0 * sy + x1 * sx);
  const float4 v110 = __ldg(b + z1 * sz + y1 * sy + x0 * sx);
  const float4 v111 = __ldg(b + z1 * sz + y1 * sy + x1 * sx);
  // ATen: tnw = (x1-x)(y1-y)(z1-z), tne = (x-x0)(y1-y)(z1-z), tsw, tse, bnw, bne, bsw, bse
  const float w000 = __fmul_rn(__fmul_rn(xl, yl), zl);
  const float w001 = xin ? __fmul_rn(__fmul_rn(xh, yl), zl) : 0.f;
  const float w010 = yin ? __fmul_rn(__fmul_rn(xl, yh), zl) : 0.f;
  const float w011 = (xin && yin) ? __fmul_rn(__fmul_rn(xh, yh), zl) : 0.f;
  const float w100 = zin ? __fmul_rn(__fmul_rn(xl, yl), zh) : 0.f;
  const float w101 = (xin && zin) ? __fmul_rn(__fmul_rn(xh, yl), zh) : 0.f;
  const float w110 = (yin && zin) ? __fmul_rn(__fmul_rn(xl, yh), zh) : 0.f;
  const float w111 = (xin && yin && zin) ? __fmul_rn(__fmul_rn(xh, yh), zh) : 0.f;
  float4 o;                                        // same fused multiply-add chain as gather_c4 (trilinear.cuh): the two gathers agree bit for bit
#define CNG_ACC(comp)                                                          \
  o.comp = __fmul_rn(v000.comp, w000);                                         \
  o.comp = fmaf(v001.comp, w001, o.comp);                                      \
  o.comp = fmaf(v010.comp, w010, o.comp);                                      \
  o.comp = fmaf(v011.comp, w011, o.comp);                                      \
  o.comp = fmaf(v100.comp, w100, o.comp);                                      \
  o.comp = fmaf(v101.comp, w101, o.comp);                                      \
  o.comp = fmaf(v110.comp, w110, o.comp);                                      \
  o.comp = fmaf(v111.comp, w111, o.comp);
  CNG_ACC(x) CNG_ACC(y) CNG_ACC(z) CNG_ACC(w)
#undef CNG_ACC
  return o;
}

struct RayParams {
  const float4* vol;        // [B, D, H, W, C/4]
  long long vol_item_stride; // in float4 units; 0 = every batch item reads item 0's volume
  int B, C4, D, H, W;
  const float* cam2world;   // [B, 16]
  const float* rays_d_cam;  // [R, 3]
  const float* t_lin;       // [S]          (coarse)
  const float* u_jitter;    // [B, R, S]    (coarse) or NULL
  const float* t_fine;      // [B, R, S]    (fine)
  int img_w, img_h, R, S;
  float4* feat;             // [B, R, S, C/4]
  float* t_out;             // [B, R, S]    (coarse)
  float* points_out;        // [B, R, S, 3] or NULL
};

// What travels from the lane that owns a point (phase A) to the 8 lanes that gather it (phase B): voxel base index with the three
// in-range flags in bits 28..30, the 8 corner weights (formed once per point, not once per lane), and the point's row in feat.
struct PointRec {
  int base_flags;
  int row;                                // ray * S + s inside the batch item, -1 for a dead lane
  float w[8];
};
__device__ __forceinline__ PointRec shfl_point(const PointRec& r, int src) {
  PointRec o;
  o.base_flags = __shfl_sync(0xffffffffu, r.base_flags, src);
  o.row = __shfl_sync(0xffffffffu, r.row, src);
#pragma unroll
  for (int i = 0; i < 8; ++i) o.w[i] = __shfl_sync(0xffffffffu, r.w[i], src);
  return o;
}

// grid = (ceil(R / 32), B); block = 8 warps.  The block owns an 8x4 pixel patch (32 rays, lane = ray in phase A);
// warp w marches the samples s = w, w + 8, ...  Phase A: one lane = one ray, position, voxel index arithmetic and the 8 corner
// weights once per point (they used to be repeated by the 8 lanes that share a point).  Phase B: 8 rounds, each serving 4 of the
// warp's 32 points with 8 lanes x float4 per point; the point record travels by warp shuffle.  32-bit offsets inside a batch item.
template <bool FINE, bool kC8>
__global__ void __launch_bounds__(256, kC8 ? CNG_K1_MIN_BLOCKS_C8 : CNG_K1_MIN_BLOCKS) raymarch_gather_kernel(RayParams p) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  int ray;
  if ((p.img_w % kTileW) == 0 && (p.img_h % kTileH) == 0) {
    const int tiles_x = p.img_w / kTileW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    ray = (ty * kTileH + lane / kTileW) * p.img_w + tx * kTileW + (lane % kTileW);
  } else {
    ray = blockIdx.x * kPointsPerBlock + lane;
  }
  const bool live = ray < p.R;
  const int rr = live ? ray : p.R - 1;                       // dead lanes shadow a valid ray and skip their stores
  const float* M = p.cam2world + 16 * b;
  const float m00 = __ldg(M + 0), m01 = __ldg(M + 1), m02 = __ldg(M + 2), m03 = __ldg(M + 3);
  const float m10 = __ldg(M + 4), m11 = __ldg(M + 5), m12 = __ldg(M + 6), m13 = __ldg(M + 7);
  const float m20 = __ldg(M + 8), m21 = __ldg(M + 9), m22 = __ldg(M + 10), m23 = __ldg(M + 11);
  const float dx = __ldg(p.rays_d_cam + 3 * rr), dy = __ldg(p.rays_d_cam + 3 * rr + 1), dz = __ldg(p.rays_d_cam + 3 * rr + 2);
  const size_t base = (static_cast<size_t>(b) * p.R + rr) * p.S;
  const float4* vol = p.vol + static_cast<size_t>(b) * p.vol_item_stride;
  float4* feat_item = p.feat ? p.feat + static_cast<size_t>(b) * p.R * p.S * p.C4 : nullptr;
  float wx = 0.f, wy = 0.f, wz = 0.f, spacing = 0.f;
  if (FINE) {
    // world-space direction: bmm(cam2world[:3,:3], d_cam)      (volumetric_rendering.py:172-180)
    wx = fmaf(m02, dz, fmaf(m01, dy, __fmul_rn(m00, dx)));
    wy = fmaf(m12, dz, fmaf(m11, dy, __fmul_rn(m10, dx)));
    wz = fmaf(m22, dz, fmaf(m21, dy, __fmul_rn(m20, dx)));
  } else {
    spacing = __fsub_rn(__ldg(p.t_lin + 1), __ldg(p.t_lin));    // z_vals[...,1] - z_vals[...,0]
  }
  const int sub = lane & 7, grp = lane >> 3;
  const char* vol_lane_bytes = reinterpret_cast<const char*>(vol + sub);                       // kC8: this lane's channel group of voxel 0
  char* feat_lane_bytes = reinterpret_cast<char*>(feat_item ? feat_item + sub : nullptr);       // ... and of the item's row 0
  for (int s = warp; s < p.S; s += 8) {
    // ---- phase A: this lane's ray, sample s ----
    float px, py, pz;
    if (FINE) {
      const float t = __ldg(p.t_fine + base + s);
      // origins + directions * t: separate mul and add          (generators.py:138-142)
      px = __fadd_rn(m03, __fmul_rn(wx, t));
      py = __fadd_rn(m13, __fmul_rn(wy, t));
      pz = __fadd_rn(m23, __fmul_rn(wz, t));
    } else {
      const float t = __ldg(p.t_lin + s);
      float cx = __fmul_rn(dx, t), cy = __fmul_rn(dy, t), cz = __fmul_rn(dz, t);   // points = d * z_vals
      float tj = t;
      if (p.u_jitter != nullptr) {
        const float off = __fmul_rn(__fsub_rn(__ldg(p.u_jitter + base + s), 0.5f), spacing);
        tj = __fadd_rn(t, off);
        cx = __fadd_rn(cx, __fmul_rn(off, dx));
        cy = __fadd_rn(cy, __fmul_rn(off, dy));
        cz = __fadd_rn(cz, __fmul_rn(off, dz));
      }
      if (live) p.t_out[base + s] = tj;
      // cam2world @ [p, 1]
      px = fmaf(m02, cz, fmaf(m01, cy, fmaf(m00, cx, m03)));
      py = fmaf(m12, cz, fmaf(m11, cy, fmaf(m10, cx, m13)));
      pz = fmaf(m22, cz, fmaf(m21, cy, fmaf(m20, cx, m23)));
    }
    if (p.points_out != nullptr && live) {
      float* o = p.points_out + 3 * (base + s);
      o[0] = px; o[1] = py; o[2] = pz;
    }
    if (p.feat == nullptr) continue;                       // points-only mode: the gather happens in the consumer (fused K2 prologue)
    const CornerRec rec = corner_record(px, py, pz, p.D, p.H, p.W);
    PointRec mine;
    mine.base_flags = rec.base | (rec.flags << 28);
    mine.row = live ? rr * p.S + s : -1;
    corner_weights(rec, mine.w);
    // ---- phase B: 4 points per round, 8 lanes x float4 each ----
#pragma unroll kK1Unroll
    for (int round = 0; round < 8; ++round) {
      const PointRec r = shfl_point(mine, round * 4 + grp);
      if constexpr (kC8) {
        // 32 channels: lane `sub` owns float4 `sub` of the point; all strides are compile-time multiples of 8 float4
        const float4 f = gather_c4_w_bytes(vol_lane_bytes, 128u, p.H, p.W, static_cast<unsigned>(r.base_flags) & 0x0fffffffu, r.base_flags >> 28, r.w);
        if (r.row >= 0) *reinterpret_cast<float4*>(feat_lane_bytes + static_cast<unsigned>(r.row) * 128u) = f;
      } else {
        for (int cg = sub; cg < p.C4; cg += kLanesPerPoint) {
          const float4 f = gather_c4_w(vol, p.H, p.W, p.C4, r.base_flags & 0x0fffffff, r.base_flags >> 28, r.w, cg);
          if (r.row >= 0) feat_item[r.row * p.C4 + cg] = f;
        }
      }
    }
  }
}

// grid = ceil(B*N / 32); one point per 8 lanes, caller-supplied positions.
__global__ void __launch_bounds__(256) gather_points_kernel(const float4* __restrict__ vol_all, int C4, int D, int H, int W,
                                                             const float* __restrict__ points, long long N, long long total,
                                                             float4* __restrict__ feat, int32_t* __restrict__ corner_idx) {
  const int sub = threadIdx.x & (kLanesPerPoint - 1);
  const long long i = static_cast<long long>(blockIdx.x) * kPointsPerBlock + threadIdx.x / kLanesPerPoint;
  if (i >= total) return;
  const long long b = i / N;
  const float4* vol = vol_all + static_cast<size_t>(b) * D * H * W * C4;
  const float px = __ldg(points + 3 * i), py = __ldg(points + 3 * i + 1), pz = __ldg(points + 3 * i + 2);
  int idx[3];
  for (int cg = sub; cg < C4; cg += kLanesPerPoint)
    feat[i * C4 + cg] = trilinear_c4(vol, D, H, W, C4, px, py, pz, cg, idx);
  if (corner_idx != nullptr && sub == 0) {
    corner_idx[3 * i] = idx[0]; corner_idx[3 * i + 1] = idx[1]; corner_idx[3 * i + 2] = idx[2];
  }
}

// NCDHW -> NDHWC: block = one run of 32 voxels x all channels, staged through shared memory so
// both the reads (32 consecutive voxels of one channel) and the writes (32 voxels x C floats,
// contiguous) are full 128-byte lines.  CT = C when known at compile time (32: shifts instead of divisions).
__device__ __forceinline__ float load_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float load_as_float(const __half* p) { return __half2float(__ldg(p)); }

// TIn: float, or __half for a volume uploaded in 16 bits (half the host->device bytes; widened here, in the pass that re-lays it)
template <int CT, typename TIn>
__global__ void __launch_bounds__(256) channels_last_kernel(const TIn* __restrict__ src, float* __restrict__ dst, int C_rt,
                                                             long long vox) {
  extern __shared__ float tile[];   // [C][33]
  const int C = CT > 0 ? CT : C_rt;
  const long long v0 = static_cast<long long>(blockIdx.x) * 32;
  const int b = blockIdx.y;
  const TIn* s = src + static_cast<size_t>(b) * C * vox;
  float* d = dst + static_cast<size_t>(b) * C * vox;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long v = v0 + lane;
#pragma unroll 4
  for (int c = w; c < C; c += 8) tile[c * 33 + lane] = v < vox ? load_as_float(s + static_cast<size_t>(c) * vox + v) : 0.f;
  __syncthreads();
  const int nv = static_cast<int>(min(32LL, vox - v0));
#pragma unroll 4
  for (int e = threadIdx.x; e < nv * C; e += 256) {
    const int vv = e / C, c = e - vv * C;
    d[(v0 + vv) * C + c] = tile[c * 33 + vv];
  }
}

static int check_volume(const void* vol, int B, int C, int D, int H, int W, const char* who) {
  CNG_REQUIRE(vol != nullptr || B == 0, CNG_ERR_INVALID_ARGUMENT, "%s: NULL volume", who);
  CNG_REQUIRE(B >= 0 && C >= 1 && D >= 1 && H >= 1 && W >= 1, CNG_ERR_INVALID_ARGUMENT, "%s: bad volume shape", who);
  CNG_REQUIRE(C % 4 == 0 && C <= 128, CNG_ERR_UNSUPPORTED, "%s: C=%d (need C %% 4 == 0 and C <= 128)", who, C);
  CNG_REQUIRE((reinterpret_cast<uintptr_t>(vol) & 15) == 0, CNG_ERR_INVALID_ARGUMENT, "%s: volume not 16-byte aligned", who);
  // the gather addresses a batch item with 32-bit BYTE offsets (items under 4 GB) and carries the voxel index in 28 bits (trilinear.cuh, PointRec)
  CNG_REQUIRE(static_cast<long long>(D) * H * W * (C / 4) < (1LL << 28), CNG_ERR_UNSUPPORTED,
              "%s: volume of %d x %d x %d x %d exceeds the gather's 32-bit offsets", who, D, H, W, C);
  return CNG_OK;
}

template <typename TIn>
static int volume_to_channels_last_impl(const TIn* vol_ncdhw, float* vol_ndhwc, int B, int C, int D, int H, int W, cng_stream_t stream) {
  CNG_REQUIRE(vol_ncdhw && vol_ndhwc, CNG_ERR_INVALID_ARGUMENT, "volume_to_channels_last: NULL pointer");
  CNG_REQUIRE(B >= 0 && C >= 1 && D >= 1 && H >= 1 && W >= 1, CNG_ERR_INVALID_ARGUMENT, "volume_to_channels_last: bad shape");
  CNG_REQUIRE(C <= 256 && B <= 65535, CNG_ERR_UNSUPPORTED, "volume_to_channels_last: C=%d B=%d", C, B);
  if (B == 0) return CNG_OK;
  if (int e = cng_device_check()) return e;
  const long long vox = static_cast<long long>(D) * H * W;
  dim3 grid(static_cast<unsigned>((vox + 31) / 32), B);
  const size_t smem = static_cast<size_t>(C) * 33 * sizeof(float);
  if (C == 32) channels_last_kernel<32, TIn><<<grid, 256, smem, as_stream(stream)>>>(vol_ncdhw, vol_ndhwc, C, vox);
  else channels_last_kernel<0, TIn><<<grid, 256, smem, as_stream(stream)>>>(vol_ncdhw, vol_ndhwc, C, vox);
  return check_launch("cng_volume_to_channels_last");
}

}  // namespace cng

extern "C" {

int cng_volume_to_channels_last(const float* vol_ncdhw, float* vol_ndhwc, int B, int C, int D, int H, int W, cng_stream_t stream) {
  return cng::volume_to_channels_last_impl<float>(vol_ncdhw, vol_ndhwc, B, C, D, H, W, stream);
}

int cng_volume_f16_to_channels_last(const void* vol_ncdhw_f16, float* vol_ndhwc, int B, int C, int D, int H, int W, cng_stream_t stream) {
  return cng::volume_to_channels_last_impl<__half>(static_cast<const __half*>(vol_ncdhw_f16), vol_ndhwc, B, C, D, H, W, stream);
}

static int raymarch_common(bool fine, const float* vol, long long vol_item_stride, int B, int C, int D, int H, int W, const float* cam2world,
                           const float* rays_d_cam, const float* t_lin, const float* u_jitter, const float* t_fine,
                           int img_w, int img_h, int S, float* feat, float* t_out, float* points_out,
                           cng_stream_t stream) {
  const char* who = fine ? "raymarch_gather_fine" : "raymarch_gather_coarse";
  if (int e = cng::check_volume(vol, B, C, D, H, W, who)) return e;
  CNG_REQUIRE(B == 0 || (cam2world && rays_d_cam && (feat || points_out)), CNG_ERR_INVALID_ARGUMENT, "%s: NULL pointer", who);
  CNG_REQUIRE(B == 0 || (fine ? (t_fine != nullptr) : (t_lin != nullptr && t_out != nullptr)), CNG_ERR_INVALID_ARGUMENT,
              "%s: NULL distance buffer", who);
  CNG_REQUIRE(img_w >= 1 && img_h >= 1 && S >= (fine ? 1 : 2), CNG_ERR_INVALID_ARGUMENT, "%s: img=%dx%d S=%d", who, img_w, img_h, S);
  CNG_REQUIRE(B <= 65535, CNG_ERR_UNSUPPORTED, "%s: B=%d > 65535", who, B);
  CNG_REQUIRE(static_cast<long long>(img_w) * img_h * S * (C / 4) < (1LL << 28), CNG_ERR_UNSUPPORTED,
              "%s: %d x %d rays x %d samples x %d channels per item exceeds the 32-bit byte offsets of feat (4 GB per batch item)", who, img_w, img_h, S, C);
  CNG_REQUIRE((reinterpret_cast<uintptr_t>(feat) & 15) == 0, CNG_ERR_INVALID_ARGUMENT, "%s: feat not 16-byte aligned", who);
  if (B == 0) return CNG_OK;
  if (int e = cng_device_check()) return e;
  cng::RayParams p{};
  CNG_REQUIRE(vol_item_stride == 0 || vol_item_stride == static_cast<long long>(C) * D * H * W, CNG_ERR_INVALID_ARGUMENT,
              "%s: vol_item_stride must be 0 (shared volume) or C*D*H*W", who);
  p.vol = reinterpret_cast<const float4*>(vol); p.vol_item_stride = vol_item_stride / 4; p.B = B; p.C4 = C / 4; p.D = D; p.H = H; p.W = W;
  p.cam2world = cam2world; p.rays_d_cam = rays_d_cam; p.t_lin = t_lin; p.u_jitter = u_jitter; p.t_fine = t_fine;
  p.img_w = img_w; p.img_h = img_h; p.R = img_w * img_h; p.S = S;
  p.feat = reinterpret_cast<float4*>(feat); p.t_out = t_out; p.points_out = points_out;
  dim3 grid((p.R + cng::kPointsPerBlock - 1) / cng::kPointsPerBlock, B);
  if (p.C4 == 8) {                                          // the shipped shape: 32 feature channels
    if (fine) cng::raymarch_gather_kernel<true, true><<<grid, 256, 0, cng::as_stream(stream)>>>(p);
    else cng::raymarch_gather_kernel<false, true><<<grid, 256, 0, cng::as_stream(stream)>>>(p);
  } else {
    if (fine) cng::raymarch_gather_kernel<true, false><<<grid, 256, 0, cng::as_stream(stream)>>>(p);
    else cng::raymarch_gather_kernel<false, false><<<grid, 256, 0, cng::as_stream(stream)>>>(p);
  }
  return cng::check_launch(who);
}

int cng_raymarch_gather_coarse(const float* vol_ndhwc, long long vol_item_stride, int B, int C, int D, int H, int W, const float* cam2world,
                               const float* rays_d_cam, const float* t_lin, const float* u_jitter, int img_w,
                               int img_h, int S, float* feat, float* t_out, float* points_out, cng_stream_t stream) {
  return raymarch_common(false, vol_ndhwc, vol_item_stride, B, C, D, H, W, cam2world, rays_d_cam, t_lin, u_jitter, nullptr, img_w, img_h,
                         S, feat, t_out, points_out, stream);
}

int cng_raymarch_gather_fine(const float* vol_ndhwc, long long vol_item_stride, int B, int C, int D, int H, int W, const float* cam2world,
                             const float* rays_d_cam, const float* t_fine, int img_w, int img_h, int S, float* feat,
                             float* points_out, cng_stream_t stream) {
  return raymarch_common(true, vol_ndhwc, vol_item_stride, B, C, D, H, W, cam2world, rays_d_cam, nullptr, nullptr, t_fine, img_w, img_h, S,
                         feat, nullptr, points_out, stream);
}

int cng_gather_points(const float* vol_ndhwc, int B, int C, int D, int H, int W, const float* points, long long N,
                      float* feat, int32_t* corner_idx, cng_stream_t stream) {
  if (int e = cng::check_volume(vol_ndhwc, B, C, D, H, W, "gather_points")) return e;
  CNG_REQUIRE(N >= 0, CNG_ERR_INVALID_ARGUMENT, "gather_points: N=%lld", N);
  CNG_REQUIRE(static_cast<long long>(B) * N == 0 || (points && feat), CNG_ERR_INVALID_ARGUMENT, "gather_points: NULL pointer");
  CNG_REQUIRE((reinterpret_cast<uintptr_t>(feat) & 15) == 0, CNG_ERR_INVALID_ARGUMENT, "gather_points: feat not 16-byte aligned");
  const long long total = static_cast<long long>(B) * N;
  if (total == 0) return CNG_OK;
  CNG_REQUIRE((total + 31) / 32 < 0x7fffffffLL, CNG_ERR_UNSUPPORTED, "gather_points: too many points");
  if (int e = cng_device_check()) return e;
  cng::gather_points_kernel<<<static_cast<unsigned>((total + 31) / 32), 256, 0, cng::as_stream(stream)>>>(
      reinterpret_cast<const float4*>(vol_ndhwc), C / 4, D, H, W, points, N, total, reinterpret_cast<float4*>(feat), corner_idx);
  return cng::check_launch("cng_gather_points");
}

}  // extern "C"
