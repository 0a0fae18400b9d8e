// Trilinear lookup pieces shared by the ray-march / gather kernels (raymarch_gather.cu) and the fused-gather prologue of the
// FiLM-SIREN kernel (film_siren_tc.cu): ATen's grid_sampler_3d index arithmetic (border padding, align_corners=False) with explicit
// round-to-nearest intrinsics, and the 8-corner accumulation of one float4 channel group.
#pragma once
#include "cng_common.cuh"

namespace cng {

__device__ __forceinline__ void axis_index(float p, int size, int& i0, float& w_lo, float& w_hi) {
  const float g = __fdiv_rn(p, 0.6f);                                 // points / (voxel_length / 2)
  float i = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), static_cast<float>(size)), 1.f), 2.f);
  i = fminf(static_cast<float>(size - 1), fmaxf(i, 0.f));             // border: clip_coordinates
  const float f = floorf(i);
  i0 = static_cast<int>(f);
  w_lo = __fsub_rn(__fadd_rn(f, 1.f), i);
  w_hi = __fsub_rn(i, f);
}

// Voxel-corner record of one sample point, computed once (phase A) and broadcast by shuffles (phase B).
struct CornerRec {
  int base;                               // ((z0 * H + y0) * W + x0), in voxels
  int flags;                              // bit0: x0+1 in range, bit1: y0+1, bit2: z0+1
  float xl, xh, yl, yh, zl, zh;
};

__device__ __forceinline__ CornerRec corner_record(float px, float py, float pz, int D, int H, int W) {
  CornerRec r;
  int x0, y0, z0;
  axis_index(px, W, x0, r.xl, r.xh);
  axis_index(py, H, y0, r.yl, r.yh);
  axis_index(pz, D, z0, r.zl, r.zh);
  r.base = (z0 * H + y0) * W + x0;
  r.flags = (x0 + 1 <= W - 1 ? 1 : 0) | (y0 + 1 <= H - 1 ? 2 : 0) | (z0 + 1 <= D - 1 ? 4 : 0);
  return r;
}

// The 8 corner weights of a point in ATen's order tnw..bse (w[z*4 + y*2 + x]); an out-of-range high corner weighs exactly 0.
__device__ __forceinline__ void corner_weights(const CornerRec& r, float (&w)[8]) {
  const bool xin = r.flags & 1, yin = r.flags & 2, zin = r.flags & 4;
  w[0] = __fmul_rn(__fmul_rn(r.xl, r.yl), r.zl);
  w[1] = xin ? __fmul_rn(__fmul_rn(r.xh, r.yl), r.zl) : 0.f;
  w[2] = yin ? __fmul_rn(__fmul_rn(r.xl, r.yh), r.zl) : 0.f;
  w[3] = (xin && yin) ? __fmul_rn(__fmul_rn(r.xh, r.yh), r.zl) : 0.f;
  w[4] = zin ? __fmul_rn(__fmul_rn(r.xl, r.yl), r.zh) : 0.f;
  w[5] = (xin && zin) ? __fmul_rn(__fmul_rn(r.xh, r.yl), r.zh) : 0.f;
  w[6] = (yin && zin) ? __fmul_rn(__fmul_rn(r.xl, r.yh), r.zh) : 0.f;
  w[7] = (xin && yin && zin) ? __fmul_rn(__fmul_rn(r.xh, r.yh), r.zh) : 0.f;
}

// 8 lanes x float4 = the 32 channels of ONE point: 8 independent 16-byte loads per lane, ATen's accumulation order, with the
// point's corner weights given (fused multiply-adds: the voxel indices and weights are bit-exact, the 8-term sum differs from
// grid_sample's separate mul / add by < 1 ulp per term -- the parity bar on features is 5e-6 abs -- at half the FP instructions).
// 32-bit element offsets: a volume item holds at most 2^31 float4 (checked by the callers).
__device__ __forceinline__ float4 gather_c4_w(const float4* __restrict__ vol, int H, int W, int C4, int base, int flags, const float (&w)[8],
                                              int cg4) {
  const int sx = (flags & 1) ? C4 : 0, sy = (flags & 2) ? W * C4 : 0, sz = (flags & 4) ? H * W * C4 : 0;
  const float4* b = vol + (base * C4 + cg4);
  const float4 v000 = __ldg(b), v001 = __ldg(b + sx), v010 = __ldg(b + sy), v011 = __ldg(b + sy + sx);
  const float4 v100 = __ldg(b + sz), v101 = __ldg(b + sz + sx), v110 = __ldg(b + sz + sy), v111 = __ldg(b + sz + sy + sx);
  float4 o;
#define CNG_ACC(comp)                                                          \
  o.comp = __fmul_rn(v000.comp, w[0]);                                         \
  o.comp = fmaf(v001.comp, w[1], o.comp);                                      \
  o.comp = fmaf(v010.comp, w[2], o.comp);                                      \
  o.comp = fmaf(v011.comp, w[3], o.comp);                                      \
  o.comp = fmaf(v100.comp, w[4], o.comp);                                      \
  o.comp = fmaf(v101.comp, w[5], o.comp);                                      \
  o.comp = fmaf(v110.comp, w[6], o.comp);                                      \
  o.comp = fmaf(v111.comp, w[7], o.comp);
  CNG_ACC(x) CNG_ACC(y) CNG_ACC(z) CNG_ACC(w)
#undef CNG_ACC
  return o;
}

// The same with the 8 corner addresses formed as ONE 64-bit base plus 32-bit BYTE offsets (nvcc otherwise carries every corner as a
// 64-bit element index and spends ~5 integer instructions per load on it: 45 of the gather kernel's 110 instructions per lane and
// round).  `vol_lane_bytes` = the lane's float4 of voxel 0 of the batch item; `row_bytes` = bytes per voxel (16 * C/4); a volume item
// is smaller than 4 GB (checked by the callers).
__device__ __forceinline__ float4 gather_c4_w_bytes(const char* __restrict__ vol_lane_bytes, unsigned row_bytes, int H, int W, unsigned base,
                                                    int flags, const float (&w)[8]) {
  const unsigned o0 = base * row_bytes;
  const unsigned ox = (flags & 1) ? row_bytes : 0u, oy = (flags & 2) ? static_cast<unsigned>(W) * row_bytes : 0u,
                 oz = (flags & 4) ? static_cast<unsigned>(H) * static_cast<unsigned>(W) * row_bytes : 0u;
  auto ld = [&](unsigned off) { return __ldg(reinterpret_cast<const float4*>(vol_lane_bytes + off)); };
  const float4 v000 = ld(o0), v001 = ld(o0 + ox), v010 = ld(o0 + oy), v011 = ld(o0 + oy + ox);
  const float4 v100 = ld(o0 + oz), v101 = ld(o0 + oz + ox), v110 = ld(o0 + oz + oy), v111 = ld(o0 + oz + oy + ox);
  float4 o;
#define CNG_ACC(comp)                                                          \
  o.comp = __fmul_rn(v000.comp, w[0]);                                         \
  o.comp = fmaf(v001.comp, w[1], o.comp);                                      \
  o.comp = fmaf(v010.comp, w[2], o.comp);                                      \
  o.comp = fmaf(v011.comp, w[3], o.comp);                                      \
  o.comp = fmaf(v100.comp, w[4], o.comp);                                      \
  o.comp = fmaf(v101.comp, w[5], o.comp);                                      \
  o.comp = fmaf(v110.comp, w[6], o.comp);                                      \
  o.comp = fmaf(v111.comp, w[7], o.comp);
  CNG_ACC(x) CNG_ACC(y) CNG_ACC(z) CNG_ACC(w)
#undef CNG_ACC
  return o;
}

// The same from a corner record (the weights are formed here: callers that serve one point with 8 lanes form them once per point
// and call gather_c4_w instead).
__device__ __forceinline__ float4 gather_c4(const float4* __restrict__ vol, int H, int W, int C4, const CornerRec& r, int cg4) {
  float w[8];
  corner_weights(r, w);
  return gather_c4_w(vol, H, W, C4, r.base, r.flags, w, cg4);
}

}  // namespace cng
