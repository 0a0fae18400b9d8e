// a14, MLP part on the 5th-gen tensor cores: the dgrad chain and the weight gradients of the FiLM-SIREN MLP as two
// tcgen05 / TMEM kernels (no library GEMM anywhere).
//
// Replaces autograd's walk back through FiLMLayer.forward x L + head (generators/siren.py:146-160, 573-579; what
// loss.backward() runs at utils.py:711) for one chunk of P points of one batch item, given the dumps of the training-mode
// forward (film_siren_tc.cu, kTrain): per layer l the output x_{l+1} = sin(u_l) as ready-made 128-point operand tile
// images, and g_l = cos(u_l) (fp16) in the epilogue's own register order.  The FiLM frequency never appears elementwise:
// with dz'_l = dy_l * cos(u_l) the chain is dy_{l-1} = dz'_l (diag(freq_l) W_l) -- the forward's folded weights, transposed --
// and the host recovers dW_l = diag(freq_l) dz'^T x, db = freq * colsum', dphase = colsum', dfreq = rowsum(W * dW') + b * colsum'.
//
//   B1  film_siren_dgrad_kernel   per 128-point tile, all layers fused, gradients never leave the SM between layers:
//         d_o = d_out (* rgb (1 - rgb))                                  prologue, split hi/lo -> A tile
//         dy_{L-1} = d_o Wf                                              one tcgen05.mma k-step (K = 16)
//         for l = L-1 .. 0:  dz_l = dy_l * g_l                           epilogue: tcgen05.ld, g streamed from HBM, bf16
//                            (dz_l tile image -> HBM, one 64 KB bulk store, for B2)
//                            dy_{l-1} = dz_l W_l                         16 x tcgen05.mma 128x256x16 (B = W_l^T images)
//         d_feat = dz_0 W_0                                              16 x tcgen05.mma 128x32x16, written fp32
//       Same organisation as the forward kernel: one persistent CTA per SM, two tiles in flight ("ping-pong": the tensor
//       pipe runs one tile's layer while the other tile's epilogue warps form dz), weights through a 3 x 32 KB ring filled
//       by cp.async.bulk, TMEM = 2 x 256 fp32 columns.
//   B2  film_siren_wgrad_kernel   dW_l += dz_l^T x_l as a split-K GEMM over the points: each CTA owns a slab of tiles, walks
//       the layers, accumulates the full 256 x 256 fp32 dW_l of its slab in TMEM (2 x 256 columns = all 512) and flushes it
//       once per layer with red.global.add.v4.f32.  Both operands are read straight from the tile images with MN-MAJOR
//       shared-memory descriptors (the point index is the contraction index), so neither dz nor x is ever transposed.
//       The idle flush warps form the column sums of dz_l (d_bias; d_phase = colsum / freq on the host) from the staged tiles.
//   head_wgrad_kernel             d_final_w += d_o^T x_L (4 x 256 outputs: plain FFMA over the tile images).
//
// HBM traffic per point and layer: g 512 B + dz 512 B (B1), dz 512 B + x 512 B (B2) on top of the 1 KB the recompute writes:
// 3 KB against 4.5 KB for the round-1 sequence (separate dz kernel + two library GEMMs per layer).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "cng_common.cuh"
#include "film_siren_tc_common.cuh"

namespace cng {

// film_siren_tc.cu
int film_siren_tc_train_launch(const float* feat, long long N, int L, const float* const* w, const float* const* b, const float* freq,
                               const float* phase, const float* final_w, const float* final_b_dev, int sigmoid_rgb, int half_operands,
                               void* workspace, size_t workspace_bytes, float* out, void* dump_x, void* dump_g, void* dump_feat,
                               cudaStream_t stream, unsigned res_save_mask, unsigned res_add_mask, float* res_scratch);
size_t film_siren_tc_workspace(int B, int L);

namespace bwdtc {

constexpr int kTileImageBytes = kATileBytes;                 // 65536: [4 K-blocks][128 rows][64 x 16 bit], 128B-swizzled
constexpr int kFeatImageBytes = kABlockBytes;                // 16384: layer-0 operand block [x_hi(32) | x_lo(32)]
// g of one tile-layer: fp16 [cc 8][q 4][i 4][lane 32] x 16 B = 65536 bytes, or 8-bit codes [cc 8][q 4][h 2][lane 32] x 16 B = 32768
__host__ __device__ constexpr int g_tile_bytes(int g_bits) { return kTileM * kHID * g_bits / 8; }

// ---- W^T images for B1 (bf16, item independent) ------------------------------------------------------------------
// [head 32 KB][layer L-1: 4 x 32 KB] ... [layer 1: 4 x 32 KB][layer 0: 16 KB]
__host__ __device__ inline size_t wt_image_bytes(int L) { return static_cast<size_t>(kChunkBytes) * (1 + 4 * (L - 1)) + 16384; }
__host__ __device__ inline size_t wt_offset(int L, int l, int c) {      // l == L -> head
  if (l == L) return 0;
  if (l > 0) return static_cast<size_t>(kChunkBytes) * (1 + 4 * (L - 1 - l) + c);
  return static_cast<size_t>(kChunkBytes) * (1 + 4 * (L - 1));
}

struct WtFoldParams {
  const float* w[16];
  const float* freq;        // [L*256] or NULL: row n of W_l is scaled by freq_l[n] (the FiLM frequency folded into the dgrad operand)
  const float* final_w;
  int L;
  uint8_t* images;
};

// one thread per (image row, 8 consecutive contraction indices)
__global__ void __launch_bounds__(256) wt_fold_kernel(WtFoldParams p) {
  const int L = p.L;
  const long long n_head = 256 * 8, n_hidden = static_cast<long long>(L - 1) * 256 * 32, n_l0 = 32 * 32;
  long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_head + n_hidden + n_l0) return;
  uint16_t v[8];
  if (e < n_head) {                                    // head: B[j][c] = Wf[c][j] for c < 4, again for 4 <= c < 8 (pairs with d_o lo)
    const int j = static_cast<int>(e >> 3), k0 = static_cast<int>(e & 7) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = k0 + i;
      v[i] = to16<false>(c < 8 ? __ldg(p.final_w + (c & 3) * kHID + j) : 0.f);
    }
    *reinterpret_cast<uint4*>(p.images + wt_offset(L, L, 0) + sw128_offset(j, k0)) = *reinterpret_cast<uint4*>(v);
    return;
  }
  e -= n_head;
  if (e < n_hidden) {                                  // layer l >= 1: B[k][n] = W_l[n][k], K-blocks of 64 n
    const int l = 1 + static_cast<int>(e / (256 * 32));
    const int r = static_cast<int>(e % (256 * 32));
    const int k = r >> 5, n0 = (r & 31) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = to16<false>(__ldg(p.w[l] + static_cast<size_t>(n0 + i) * kHID + k) * (p.freq ? __ldg(p.freq + l * kHID + n0 + i) : 1.f));
    *reinterpret_cast<uint4*>(p.images + wt_offset(L, l, n0 >> 6) + sw128_offset(k, n0 & 63)) = *reinterpret_cast<uint4*>(v);
    return;
  }
  e -= n_hidden;
  {                                                    // layer 0: B[k < 32][n] = W_0[n][k], 4 K-blocks of [32 rows][64 n] (4 KB each)
    const int k = static_cast<int>(e >> 5), n0 = static_cast<int>(e & 31) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = to16<false>(__ldg(p.w[0] + static_cast<size_t>(n0 + i) * kC0 + k) * (p.freq ? __ldg(p.freq + n0 + i) : 1.f));
    *reinterpret_cast<uint4*>(p.images + wt_offset(L, 0, 0) + (n0 >> 6) * 4096 + sw128_offset(k, n0 & 63)) = *reinterpret_cast<uint4*>(v);
  }
}

// =====================================================================================================================
// B1: dgrad chain
// =====================================================================================================================
constexpr int kRingB = 3;
constexpr int kEpiWarps = 8;                              // per tile slot
constexpr int kMmaWarpB = 2 * kEpiWarps, kProducerWarpB = kMmaWarpB + 1, kThreadsB = 32 * (kProducerWarpB + 1);
constexpr uint32_t kSmemA_B = 0;
constexpr uint32_t kSmemW_B = 2 * kATileBytes;
constexpr uint32_t kSmemBar_B = kSmemW_B + kRingB * kChunkBytes;
constexpr uint32_t kSmemTotal_B = kSmemBar_B + 128;

struct DgradParams {
  const float* d_out;        // [P, 4]
  const float* out;          // [P, 4] forward output (read when sigmoid_rgb)
  int sigmoid_rgb;
  long long P, T;            // points, tiles
  int L;
  const uint8_t* wt;         // W^T images
  const uint8_t* g;          // [L][g_stride tiles][g_tile_bytes]: layer l of this call's tiles starts at g + l * g_stride * g_tile_bytes
  int g_bits;                // 16 (fp16) or 8 (codes round(127 cos) + 128)
  long long g_stride;        // >= T (a dump that holds more tiles than this call processes, e.g. all items of a batch)
  uint8_t* dz;               // [L][T][65536] tile images (bf16), written
  float* d_feat;             // [P, 32]
  float* d_final_b;          // [4], accumulated
  // residual blocks: the activation kept by a save layer also receives the adding layer's dz (masks as in the forward)
  uint32_t res_save_mask, res_add_mask;
  float* res_scratch;        // per CTA and slot [64 column quads][128 rows] float4
};

template <bool kRes, bool kG8>
__global__ void __launch_bounds__(kThreadsB, 1) film_siren_dgrad_kernel(DgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  const uint32_t bar0 = s_base + kSmemBar_B;
  auto w_full = [&](int s) { return bar0 + 8u * s; };
  auto w_empty = [&](int s) { return bar0 + 24u + 8u * s; };
  auto act_ready = [&](int x) { return bar0 + 48u + 8u * x; };
  auto acc_full = [&](int x) { return bar0 + 64u + 8u * x; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar_B + 96);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRingB; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int x = 0; x < 2; ++x) { mbar_init(act_ready(x), 32 * kEpiWarps); mbar_init(acc_full(x), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarpB) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_base + kSmemBar_B + 96), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const long long G = gridDim.x, first = blockIdx.x;

  // step s = 0: head (1 chunk); s = 1 .. L-1: layer l = L - s (4 chunks); s = L: layer 0 (1 chunk of 16 KB)
  if (warp == kProducerWarpB) {
    const bool elected = elect_one();
    int slot = 0;
    uint32_t phase = 0;
    for (long long t0 = first; t0 < p.T; t0 += 2 * G) {
      const int nx = (t0 + G < p.T) ? 2 : 1;
      for (int s = 0; s <= L; ++s) {
        const int l = (s == 0) ? L : L - s;
        const int nchunks = (s == 0 || s == L) ? 1 : 4;
        const uint32_t bytes = (s == L) ? 16384u : static_cast<uint32_t>(kChunkBytes);
        for (int x = 0; x < nx; ++x)
          for (int c = 0; c < nchunks; ++c) {
            mbar_wait(w_empty(slot), phase ^ 1);
            if (elected) {
              mbar_arrive_expect_tx(w_full(slot), bytes);
              bulk_g2s_hint(s_base + kSmemW_B + slot * kChunkBytes, p.wt + wt_offset(L, l, c), bytes, w_full(slot), kL2EvictLast);
            }
            __syncwarp();
            if (++slot == kRingB) { slot = 0; phase ^= 1; }
          }
      }
    }
  } else if (warp == kMmaWarpB) {
    const bool elected = elect_one();
    int slot = 0;
    uint32_t phase = 0, act_phase = 0;
    constexpr uint32_t idesc_main = make_idesc(128, 256, false);
    constexpr uint32_t idesc_l0 = make_idesc(128, 32, false);
    for (long long t0 = first; t0 < p.T; t0 += 2 * G) {
      const int nx = (t0 + G < p.T) ? 2 : 1;
      for (int s = 0; s <= L; ++s) {
        const int nchunks = (s == 0 || s == L) ? 1 : 4;
        for (int x = 0; x < nx; ++x) {
          mbar_wait_lean(act_ready(x), (act_phase >> x) & 1u);
          act_phase ^= 1u << x;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(x) * kHID;
          const uint64_t a_desc0 = make_desc(s_base + kSmemA_B + x * kATileBytes);
          for (int c = 0; c < nchunks; ++c) {
            mbar_wait_lean(w_full(slot), phase);
            tc_fence_after();
            const uint64_t b_desc = make_desc(s_base + kSmemW_B + slot * kChunkBytes);
            if (elected) {
              if (s == 0) {
                tc_mma_bf16(d_tmem, a_desc0, b_desc, idesc_main, 0u);                    // K = 16: [d_o hi | d_o lo | 0]
              } else if (s < L) {
                const uint64_t a_desc = a_desc0 + c * (kABlockBytes >> 4);
                tc_mma_bf16(d_tmem, a_desc, b_desc, idesc_main, c ? 1u : 0u);
                tc_mma_bf16(d_tmem, a_desc + 2, b_desc + 2, idesc_main, 1u);
                tc_mma_bf16(d_tmem, a_desc + 4, b_desc + 4, idesc_main, 1u);
                tc_mma_bf16(d_tmem, a_desc + 6, b_desc + 6, idesc_main, 1u);
              } else {
#pragma unroll
                for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    tc_mma_bf16(d_tmem, a_desc0 + kb * (kABlockBytes >> 4) + 2 * ks, b_desc + kb * (4096 >> 4) + 2 * ks, idesc_l0,
                                (kb | ks) ? 1u : 0u);
              }
              tc_commit(w_empty(slot));
            }
            __syncwarp();
            if (++slot == kRingB) { slot = 0; phase ^= 1; }
          }
          if (elected) tc_commit(acc_full(x));
          __syncwarp();
        }
      }
    }
  } else {
    // =========================== epilogue warps (slot x = warp / 8) ===========================
    const int x = warp / kEpiWarps;
    const int wl = warp % kEpiWarps;
    const int q = warp & 3;                         // TMEM lane quarter == warp_id % 4
    const int half = wl >> 2;                       // column group: blocks [4*half, 4*half + 4) of 32 columns
    const int row = q * 32 + lane;
    const uint32_t a_base = kSmemA_B + x * kATileBytes;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(x) * kHID;
    constexpr int kSlotThreads = 32 * kEpiWarps;
    const bool storer = wl == 0 && lane == 0;        // issues (and waits for) the slot's bulk stores
    uint32_t acc_phase = 0;
    // before the A tile is written again: the previous bulk store has finished READING it
    auto tile_free = [&]() {
      if (storer) bulk_wait_read_all();
      named_bar_sync(1 + x, kSlotThreads);
    };
    for (long long t = first + x * G; t < p.T; t += 2 * G) {
      const long long n0 = t * kTileM;
      const int rows = static_cast<int>(min(static_cast<long long>(kTileM), p.P - n0));
      // ---- prologue: d_o -> A block 0, columns [hi(4) | lo(4) | 0(8)] ----
      tile_free();
      if (half == 0) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows) {
          d = __ldg(reinterpret_cast<const float4*>(p.d_out) + n0 + row);
          if (p.sigmoid_rgb) {
            const float4 y = __ldg(reinterpret_cast<const float4*>(p.out) + n0 + row);
            d.x *= y.x * (1.f - y.x);
            d.y *= y.y * (1.f - y.y);
            d.z *= y.z * (1.f - y.z);
          }
        }
        const float hx = from16<false>(to16<false>(d.x)), hy = from16<false>(to16<false>(d.y)), hz = from16<false>(to16<false>(d.z)),
                    hw = from16<false>(to16<false>(d.w));
        uint4 c0;
        c0.x = pack2<false>(d.x, d.y); c0.y = pack2<false>(d.z, d.w);
        c0.z = pack2<false>(d.x - hx, d.y - hy); c0.w = pack2<false>(d.z - hz, d.w - hw);
        uint8_t* blk = smem + a_base + row * 128;
        *reinterpret_cast<uint4*>(blk + ((0 ^ (row & 7)) << 4)) = c0;
        *reinterpret_cast<uint4*>(blk + ((1 ^ (row & 7)) << 4)) = make_uint4(0, 0, 0, 0);
        // d_final_b += column sums of d_o
        const float sx = warp_sum(d.x), sy = warp_sum(d.y), sz = warp_sum(d.z), sw = warp_sum(d.w);
        if (lane == 0) {
          atomicAdd(p.d_final_b + 0, sx); atomicAdd(p.d_final_b + 1, sy); atomicAdd(p.d_final_b + 2, sz); atomicAdd(p.d_final_b + 3, sw);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(act_ready(x));
      for (int s = 0; s < L; ++s) {
        const int l = L - 1 - s;                                           // this epilogue forms dz_l from dy_l (accumulator) and g_l
        constexpr bool g8 = kG8;
        const uint4* gt = reinterpret_cast<const uint4*>(p.g + (static_cast<size_t>(l) * p.g_stride + t) * g_tile_bytes(g8 ? 8 : 16));
        uint4 ga[4], gb[4];
        auto load_g = [&](uint4 (&gg)[4], int cc) {
          if constexpr (g8) {
#pragma unroll
            for (int h = 0; h < 2; ++h) gg[h] = ld_global_evict_first(gt + ((cc * 4 + q) * 2 + h) * 32 + lane);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) gg[i] = ld_global_evict_first(gt + ((cc * 4 + q) * 4 + i) * 32 + lane);
          }
        };
        load_g(ga, 4 * half);
        mbar_wait(acc_full(x), acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        tile_free();
        float4* rsd = nullptr;
        if constexpr (kRes) rsd = reinterpret_cast<float4*>(p.res_scratch) + (static_cast<size_t>(blockIdx.x) * 2 + x) * (kHID / 4) * kTileM;
        auto finish_block = [&](const uint4 (&gg)[4], int cc) {
          uint32_t v[32];
          CNG_TMEM_LD_32(t_lane + cc * 32, v);
          tmem_ld_wait();
          if constexpr (kRes) {
            // output of layer l is a kept activation: add the dz that the adding layer left for it
            if (((p.res_save_mask >> l) & 1u) && (p.res_add_mask >> (l + 1))) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 r4 = rsd[static_cast<size_t>(cc * 8 + i) * kTileM + row];
                v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + r4.x);
                v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + r4.y);
                v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + r4.z);
                v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + r4.w);
              }
            }
          }
          uint8_t* blk = smem + a_base + (cc >> 1) * kABlockBytes + row * 128;
          float dzv[32];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            // columns 8 i .. 8 i + 7 of the block.  fp16: the four words of piece i.  8-bit: two words of codes u = round(127 cos) + 128;
            // a code is placed in mantissa bits 8..15 of 1.0f (one PRMT: 1 + u 2^-15) and scaled back with one FMA:
            // (f - 1 - 2^-8) 2^15 / 127 = (u - 128) / 127
            const uint32_t gw16[4] = {gg[i].x, gg[i].y, gg[i].z, gg[i].w};
            const uint32_t gw8[2] = {(i & 1) ? gg[i >> 1].z : gg[i >> 1].x, (i & 1) ? gg[i >> 1].w : gg[i >> 1].y};
            constexpr float kGs = 32768.f / 127.f, kGo = -(1.f + 1.f / 256.f) * (32768.f / 127.f);
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float2 gf;
              if constexpr (g8) {
                gf.x = fmaf(__uint_as_float(__byte_perm(gw8[j >> 1], 0x3F800000u, (j & 1) ? 0x7624 : 0x7604)), kGs, kGo);
                gf.y = fmaf(__uint_as_float(__byte_perm(gw8[j >> 1], 0x3F800000u, (j & 1) ? 0x7634 : 0x7614)), kGs, kGo);
              } else {
                gf = __half22float2(*reinterpret_cast<const __half2*>(&gw16[j]));
              }
              const float a = __uint_as_float(v[8 * i + 2 * j]) * gf.x, b = __uint_as_float(v[8 * i + 2 * j + 1]) * gf.y;
              dzv[8 * i + 2 * j] = a; dzv[8 * i + 2 * j + 1] = b;
              o[j] = pack2<false>(a, b);
            }
            const int chunk = ((cc & 1) * 4 + i) ^ (row & 7);
            *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(o[0], o[1], o[2], o[3]);
          }
          if constexpr (kRes) {
            // layer l adds a kept activation to its pre-activation: that activation's gradient receives dz_l
            if ((p.res_add_mask >> l) & 1u) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                rsd[static_cast<size_t>(cc * 8 + i) * kTileM + row] = make_float4(dzv[4 * i], dzv[4 * i + 1], dzv[4 * i + 2], dzv[4 * i + 3]);
            }
          }
        };
#pragma unroll
        for (int i = 0; i < 4; i += 2) {
          const int cc = 4 * half + i;
          load_g(gb, cc + 1);
          finish_block(ga, cc);
          if (i + 2 < 4) load_g(ga, cc + 2);
          finish_block(gb, cc + 1);
        }
        tc_fence_before();
        fence_proxy_async();
        named_bar_sync(1 + x, kSlotThreads);              // the whole dz_l tile is in shared memory (and fenced for the async proxy)
        if (storer) {
          bulk_s2g_hint(p.dz + (static_cast<size_t>(l) * p.T + t) * kTileImageBytes, s_base + a_base, kTileImageBytes, kL2EvictFirst);
          bulk_commit();
        }
        mbar_arrive(act_ready(x));
      }
      // ---- d_feat = dz_0 W_0: 32 accumulator columns (the column-half-0 warps hold them) ----
      mbar_wait(acc_full(x), acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      if (half == 0) {
        uint32_t v[32];
        CNG_TMEM_LD_32(t_lane, v);
        tmem_ld_wait();
        if (row < rows) {
          float4* dst = reinterpret_cast<float4*>(p.d_feat + (n0 + row) * kC0);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
      }
      tc_fence_before();
    }
    if (storer) bulk_wait_all();                      // the last stores have left shared memory before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarpB) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// =====================================================================================================================
// B2: weight gradients, split-K over the points
// =====================================================================================================================
constexpr int kStageBytes = 65536;                        // dz half tile (64 points x 256) 32 KB + x half tile 32 KB
constexpr int kStages = 3;
constexpr uint32_t kSmemBar_W = kStages * kStageBytes;    // 196608
constexpr uint32_t kSmemTotal_W = kSmemBar_W + 128;
constexpr int kFlushWarps = 16, kMmaWarpW = 16, kProducerWarpW = 17, kThreadsW = 32 * 18;

struct WgradParams {
  const uint8_t* dz;         // [L][T][65536] (bf16)
  const uint8_t* x;          // [L][x_stride tiles][65536]: x[l] = output of layer l
  const uint8_t* feat;       // [T][16384]: layer-0 operand blocks [x_hi | x_lo]
  long long T, x_stride;
  int L;
  int x_half;                // 1: x / feat hold fp16, 0: bf16
  float* dW[16];             // [256][K_l] fp32, accumulated
  float* colsum;             // [L][256], accumulated
};
static_assert(kSmemBar_W + 128 <= 232448, "wgrad shared memory");

// MN-major SWIZZLE_128B descriptor: 64 MN elements (128 B) x 8 contraction rows per 1 KB atom; next 64 MN elements at
// `lbo` bytes, next 8 contraction rows at 1024 bytes (cute::UMMA make_umma_desc<Major::MN>, LayoutType::B128).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_mn(int M, int N, uint32_t a_fmt, uint32_t b_fmt) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | (1u << 15) | (1u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreadsW, 1) film_siren_wgrad_kernel(WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  const uint32_t bar0 = s_base + kSmemBar_W;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 24u + 8u * s; };
  const uint32_t acc_full = bar0 + 48u, acc_empty = bar0 + 56u;
  auto ready = [&](int s) { return bar0 + 64u + 8u * s; };      // x operand converted to bf16 (fp16 dumps only)
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar_W + 96);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1 + kFlushWarps); mbar_init(ready(s), kFlushWarps); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, kFlushWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarpW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_base + kSmemBar_W + 96), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const long long G = gridDim.x;
  const long long t_begin = blockIdx.x * p.T / G, t_end = (blockIdx.x + 1) * p.T / G;
  const long long n_stage = 2 * (t_end - t_begin);             // half tiles per layer

  if (warp == kProducerWarpW) {
    const bool elected = elect_one();
    int slot = 0;
    uint32_t phase = 0;
    for (int l = 0; l < L; ++l)
      for (long long t = t_begin; t < t_end; ++t)
        for (int hf = 0; hf < 2; ++hf) {
          mbar_wait(empty(slot), phase ^ 1);
          if (elected) {
            const uint32_t st = s_base + slot * kStageBytes;
            mbar_arrive_expect_tx(full(slot), 32768u + (l == 0 ? 8192u : 32768u));
            const uint8_t* dz = p.dz + (static_cast<size_t>(l) * p.T + t) * kTileImageBytes + hf * 8192;
#pragma unroll
            for (int j = 0; j < 4; ++j) bulk_g2s_hint(st + j * 8192, dz + j * kABlockBytes, 8192, full(slot), kL2EvictFirst);
            if (l == 0) {
              bulk_g2s_hint(st + 32768, p.feat + static_cast<size_t>(t) * kFeatImageBytes + hf * 8192, 8192, full(slot), kL2EvictFirst);
            } else {
              const uint8_t* xs = p.x + (static_cast<size_t>(l - 1) * p.x_stride + t) * kTileImageBytes + hf * 8192;
#pragma unroll
              for (int j = 0; j < 4; ++j) bulk_g2s_hint(st + 32768 + j * 8192, xs + j * kABlockBytes, 8192, full(slot), kL2EvictFirst);
            }
          }
          __syncwarp();
          if (++slot == kStages) { slot = 0; phase ^= 1; }
        }
  } else if (warp == kMmaWarpW) {
    const bool elected = elect_one();
    int slot = 0;
    uint32_t phase = 0;
    // both operands bf16: tcgen05.mma kind::f16 traps (illegal instruction, measured) when A is bf16 and B fp16, so an fp16 x
    // dump is converted to bf16 in shared memory by the flush warps before the MMA reads it (`ready` barriers)
    const uint32_t idesc_main = make_idesc_mn(128, 256, 1u, 1u), idesc_l0 = make_idesc_mn(128, 64, 1u, 1u);
    const bool convert = p.x_half != 0;
    for (int l = 0; l < L; ++l) {
      if (l > 0) {                                           // the flush warps have read layer l-1's accumulator
        mbar_wait_lean(acc_empty, static_cast<uint32_t>(l - 1) & 1u);
        tc_fence_after();
      }
      const uint32_t idesc = l == 0 ? idesc_l0 : idesc_main;
      for (long long i = 0; i < n_stage; ++i) {
        mbar_wait_lean(convert ? ready(slot) : full(slot), phase);
        tc_fence_after();
        const uint32_t st = s_base + slot * kStageBytes;
        if (elected) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              tc_mma_bf16(tmem_base + h * 256, make_desc_mn(st + h * 16384 + ks * 2048, 8192), make_desc_mn(st + 32768 + ks * 2048, 8192), idesc,
                          (i > 0 || ks > 0) ? 1u : 0u);
          tc_commit(empty(slot));
        }
        __syncwarp();
        if (++slot == kStages) { slot = 0; phase ^= 1; }
      }
      if (elected) tc_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ---- flush warps: column sums of dz while the layer streams, then TMEM -> red.global.add ----
    const int tid = threadIdx.x;                       // 0..511
    const int pr = tid & 127, rg = tid >> 7;            // column pair (2 pr, 2 pr + 1), rows [16 rg, 16 rg + 16) of the half tile
    const uint32_t col_off = (pr >> 5) * 8192;
    const int cchunk = (2 * (pr & 31)) >> 3, cbyte = ((2 * (pr & 31)) & 7) * 2;
    const int q = warp & 3, cg = warp >> 2, h = cg >> 1, ch = cg & 1;
    int slot = 0;
    uint32_t phase = 0;
    for (int l = 0; l < L; ++l) {
      float c0 = 0.f, c1 = 0.f;
      for (long long i = 0; i < n_stage; ++i) {
        mbar_wait(full(slot), phase);
        const uint8_t* st = smem + slot * kStageBytes + col_off;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int rr = rg * 16 + r;
          const uint32_t w = *reinterpret_cast<const uint32_t*>(st + rr * 128 + ((cchunk ^ (rr & 7)) << 4) + cbyte);
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
          c0 += f.x;
          c1 += f.y;
        }
        if (p.x_half) {                                  // x half tile fp16 -> bf16 in place (elementwise: the swizzle is untouched)
          uint4* xq = reinterpret_cast<uint4*>(smem + slot * kStageBytes + 32768);
          const int n16 = (l == 0) ? 1 : 4;                // 8 KB (layer-0 operand block) or 32 KB: 16-byte pieces per thread
#pragma unroll 1
          for (int k = 0; k < n16; ++k) {
            uint4 w4 = xq[tid + 512 * k];
            uint32_t* w = reinterpret_cast<uint32_t*>(&w4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
              w[e] = pack2<false>(f.x, f.y);
            }
            xq[tid + 512 * k] = w4;
          }
          fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) {
          if (p.x_half) mbar_arrive(ready(slot));
          mbar_arrive(empty(slot));
        }
        if (++slot == kStages) { slot = 0; phase ^= 1; }
      }
      atomicAdd(p.colsum + l * kHID + 2 * pr, c0);
      atomicAdd(p.colsum + l * kHID + 2 * pr + 1, c1);
      mbar_wait(acc_full, static_cast<uint32_t>(l) & 1u);
      tc_fence_after();
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * 256;
      const int n = h * 128 + q * 32 + lane;
      if (l > 0) {
        float* dst = p.dW[l] + static_cast<size_t>(n) * kHID + ch * 128;
#pragma unroll 1
        for (int blk = 0; blk < 4; ++blk) {
          uint32_t v[32];
          CNG_TMEM_LD_32(t_lane + ch * 128 + blk * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            red_add_v4(dst + blk * 32 + 4 * i, __uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                       __uint_as_float(v[4 * i + 3]));
        }
      } else if (ch == 0) {                              // layer 0: dW_0[n][k] = D[n][k] (x_hi) + D[n][32 + k] (x_lo)
        uint32_t a[32], b[32];
        CNG_TMEM_LD_32(t_lane, a);
        CNG_TMEM_LD_32(t_lane + 32, b);
        tmem_ld_wait();
        float* dst = p.dW[0] + static_cast<size_t>(n) * kC0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          red_add_v4(dst + 4 * i, __uint_as_float(a[4 * i]) + __uint_as_float(b[4 * i]), __uint_as_float(a[4 * i + 1]) + __uint_as_float(b[4 * i + 1]),
                     __uint_as_float(a[4 * i + 2]) + __uint_as_float(b[4 * i + 2]), __uint_as_float(a[4 * i + 3]) + __uint_as_float(b[4 * i + 3]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarpW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ---- head: d_final_w[c][j] += sum_p d_o[p][c] x_L[p][j] ---------------------------------------------------------------
// One block per slab of tiles; x_L is read from its tile images 16 bytes at a time: lane = (K-block, logical 16-byte chunk) of
// one row, so a warp reads the four 128-byte lines of a row per load; warp w takes rows w, w + 8, ... of every tile.
template <bool kHalf>
__global__ void __launch_bounds__(256, 3) head_wgrad_kernel(const float* __restrict__ d_out, const float* __restrict__ out, int sigmoid_rgb,
                                                         const uint8_t* __restrict__ xL, long long P, long long T, float* __restrict__ d_final_w) {
  __shared__ float4 s_do[kTileM];
  __shared__ float s_red[8][4][kHID];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long t_begin = blockIdx.x * T / gridDim.x, t_end = (blockIdx.x + 1) * T / gridDim.x;
  float acc[4][8];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[c][e] = 0.f;
  const int blk = lane >> 3, ch = lane & 7;               // columns 64 blk + 8 ch .. + 7
  for (long long t = t_begin; t < t_end; ++t) {
    const long long n0 = t * kTileM;
    __syncthreads();
    if (tid < kTileM) {
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + tid < P) {
        d = __ldg(reinterpret_cast<const float4*>(d_out) + n0 + tid);
        if (sigmoid_rgb) {
          const float4 y = __ldg(reinterpret_cast<const float4*>(out) + n0 + tid);
          d.x *= y.x * (1.f - y.x);
          d.y *= y.y * (1.f - y.y);
          d.z *= y.z * (1.f - y.z);
        }
      }
      s_do[tid] = d;
    }
    __syncthreads();
    const uint8_t* img = xL + static_cast<size_t>(t) * kTileImageBytes + blk * kABlockBytes;
#pragma unroll 8
    for (int r = w; r < kTileM; r += 8) {
      const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(img + r * 128 + ((ch ^ (r & 7)) << 4)));
      const uint32_t qw[4] = {q4.x, q4.y, q4.z, q4.w};
      const float4 d = s_do[r];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float x0, x1;
        if (kHalf) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&qw[e]));
          x0 = f.x; x1 = f.y;
        } else {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&qw[e]));
          x0 = f.x; x1 = f.y;
        }
        acc[0][2 * e] = fmaf(d.x, x0, acc[0][2 * e]); acc[0][2 * e + 1] = fmaf(d.x, x1, acc[0][2 * e + 1]);
        acc[1][2 * e] = fmaf(d.y, x0, acc[1][2 * e]); acc[1][2 * e + 1] = fmaf(d.y, x1, acc[1][2 * e + 1]);
        acc[2][2 * e] = fmaf(d.z, x0, acc[2][2 * e]); acc[2][2 * e + 1] = fmaf(d.z, x1, acc[2][2 * e + 1]);
        acc[3][2 * e] = fmaf(d.w, x0, acc[3][2 * e]); acc[3][2 * e + 1] = fmaf(d.w, x1, acc[3][2 * e + 1]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int e = 0; e < 8; ++e) s_red[w][c][blk * 64 + ch * 8 + e] = acc[c][e];
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float a = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) a += s_red[ww][c][tid];
    atomicAdd(d_final_w + c * kHID + tid, a);
  }
}

struct Layout {
  size_t fold, wt, xs, gs, dzs, feat, out, total;
};
static Layout layout(long long P, int L) {
  auto up = [](size_t v) { return (v + 1023) & ~static_cast<size_t>(1023); };
  const size_t T = static_cast<size_t>((P + kTileM - 1) / kTileM);
  Layout l{};
  size_t off = 0;
  l.fold = off; off += up(film_siren_tc_workspace(1, L));
  l.wt = off; off += up(wt_image_bytes(L));
  l.xs = off; off += up(static_cast<size_t>(L) * T * kTileImageBytes);
  l.gs = off; off += up(static_cast<size_t>(L) * T * g_tile_bytes(16));      // sized for the larger format
  l.dzs = off; off += up(static_cast<size_t>(L) * T * kTileImageBytes);
  l.feat = off; off += up(T * kFeatImageBytes);
  l.out = off; off += up(static_cast<size_t>(P) * 16);
  l.total = off;
  return l;
}

static int set_smem(const void* fn, uint32_t bytes, bool (&cache)[64], const char* what) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fail(CNG_ERR_NO_DEVICE, "%s: no current device", what);
  if (!cache[dev]) {
    const cudaError_t ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "%s: smem attribute: %s", what, cudaGetErrorString(ce));
    cache[dev] = true;
  }
  return CNG_OK;
}

int dgrad_launch(const DgradParams& p, cudaStream_t st) {
  static bool cached[4][64] = {};
  const bool res = (p.res_save_mask | p.res_add_mask) != 0, g8 = p.g_bits == 8;
  void (*fn)(DgradParams) = res ? (g8 ? film_siren_dgrad_kernel<true, true> : film_siren_dgrad_kernel<true, false>)
                                : (g8 ? film_siren_dgrad_kernel<false, true> : film_siren_dgrad_kernel<false, false>);
  if (int e = set_smem(reinterpret_cast<const void*>(fn), kSmemTotal_B, cached[(res ? 2 : 0) + (g8 ? 1 : 0)], "film_siren_dgrad")) return e;
  const unsigned grid = static_cast<unsigned>(min(static_cast<long long>(sm_count()), p.T));
  fn<<<grid, kThreadsB, kSmemTotal_B, st>>>(p);
  return check_launch("cng_film_siren_dgrad");
}

int wgrad_launch(const WgradParams& p, cudaStream_t st) {
  static bool c[64] = {};
  if (int e = set_smem(reinterpret_cast<const void*>(film_siren_wgrad_kernel), kSmemTotal_W, c, "film_siren_wgrad")) return e;
  const unsigned grid = static_cast<unsigned>(min(static_cast<long long>(sm_count()), p.T));
  film_siren_wgrad_kernel<<<grid, kThreadsW, kSmemTotal_W, st>>>(p);
  return check_launch("cng_film_siren_wgrad");
}

}  // namespace bwdtc
}  // namespace cng

extern "C" {

size_t cng_film_siren_bwd_workspace_bytes(long long P, int C, int HID, int L) {
  if (P <= 0 || C != 32 || HID != 256 || L < 1 || L > 16) return 0;
  return cng::bwdtc::layout(P, L).total;
}

int cng_film_siren_g_dump_bits(void) { return cng::film_siren_g_dump_bits(); }

size_t cng_film_siren_wt_image_bytes(int L) { return (L < 1 || L > 16) ? 0 : cng::bwdtc::wt_image_bytes(L); }

int cng_film_siren_wt_images(const float* const* layer_w_host, const float* freq, const float* final_w, int C, int HID, int L, void* images,
                             cng_stream_t stream) {
  using namespace cng;
  using namespace cng::bwdtc;
  CNG_REQUIRE(C == kC0 && HID == kHID && L >= 1 && L <= 16, CNG_ERR_UNSUPPORTED, "film_siren_wt_images: needs C=32, HID=256, L<=16");
  CNG_REQUIRE(layer_w_host && final_w && images && (reinterpret_cast<uintptr_t>(images) & 15) == 0, CNG_ERR_INVALID_ARGUMENT,
              "film_siren_wt_images: NULL pointer or images not 16-byte aligned");
  if (int e = cng_device_check()) return e;
  WtFoldParams fp{};
  for (int l = 0; l < L; ++l) {
    CNG_REQUIRE(layer_w_host[l], CNG_ERR_INVALID_ARGUMENT, "film_siren_wt_images: NULL layer %d", l);
    fp.w[l] = layer_w_host[l];
  }
  fp.freq = freq; fp.final_w = final_w; fp.L = L; fp.images = static_cast<uint8_t*>(images);
  const long long n = 256 * 8 + static_cast<long long>(L - 1) * 256 * 32 + 32 * 32;
  wt_fold_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream)>>>(fp);
  return check_launch("cng_film_siren_wt_images");
}

int cng_film_siren_dgrad(const float* d_out, const float* out, int sigmoid_rgb, long long P, int L, const void* wt_images,
                         const void* g_dump, long long g_layer_stride_tiles, void* dz_dump, float* d_feat, float* d_final_b_acc,
                         unsigned res_save_mask, unsigned res_add_mask, void* res_scratch, size_t res_scratch_bytes, cng_stream_t stream) {
  using namespace cng;
  using namespace cng::bwdtc;
  CNG_REQUIRE(P >= 0 && L >= 1 && L <= 16, CNG_ERR_INVALID_ARGUMENT, "film_siren_dgrad: bad shape");
  if (P == 0) return CNG_OK;
  CNG_REQUIRE(d_out && wt_images && g_dump && dz_dump && d_feat && d_final_b_acc && (!sigmoid_rgb || out), CNG_ERR_INVALID_ARGUMENT,
              "film_siren_dgrad: NULL pointer");
  CNG_REQUIRE(((reinterpret_cast<uintptr_t>(wt_images) | reinterpret_cast<uintptr_t>(g_dump) | reinterpret_cast<uintptr_t>(dz_dump) |
                reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(d_feat)) & 15) == 0,
              CNG_ERR_INVALID_ARGUMENT, "film_siren_dgrad: buffers not 16-byte aligned");
  const bool res = (res_save_mask | res_add_mask) != 0;
  CNG_REQUIRE(((res_save_mask | res_add_mask) >> L) == 0, CNG_ERR_INVALID_ARGUMENT, "film_siren_dgrad: residual mask bit beyond layer %d", L - 1);
  CNG_REQUIRE(!res || (res_scratch && res_scratch_bytes >= cng_film_siren_res_scratch_bytes()), CNG_ERR_WORKSPACE,
              "film_siren_dgrad: residual scratch %zu < %zu bytes", res_scratch_bytes, cng_film_siren_res_scratch_bytes());
  if (res) {
    // the per-tile scratch holds ONE pending skip gradient: every adding layer must sit exactly two layers after its kept
    // layer and each kept activation must feed a single add (the shape of every ResSirenBlock chain, siren.py:218-230)
    int kept = -1;
    bool pending = false;
    for (int l = 0; l < L; ++l) {
      if ((res_add_mask >> l) & 1u) {
        CNG_REQUIRE(kept >= 0 && l - kept == 2 && pending, CNG_ERR_UNSUPPORTED,
                    "film_siren_dgrad: layer %d adds a residual that was not kept exactly two layers earlier (unsupported mask pattern)", l);
        pending = false;
      }
      if ((res_save_mask >> l) & 1u) { kept = l; pending = true; }
    }
  }
  if (int e = cng_device_check()) return e;
  DgradParams dp{};
  dp.d_out = d_out; dp.out = out; dp.sigmoid_rgb = sigmoid_rgb; dp.P = P; dp.T = (P + kTileM - 1) / kTileM; dp.L = L;
  dp.wt = static_cast<const uint8_t*>(wt_images); dp.g = static_cast<const uint8_t*>(g_dump); dp.dz = static_cast<uint8_t*>(dz_dump);
  dp.g_bits = film_siren_g_dump_bits();
  dp.g_stride = g_layer_stride_tiles > 0 ? g_layer_stride_tiles : dp.T;
  CNG_REQUIRE(dp.g_stride >= dp.T, CNG_ERR_INVALID_ARGUMENT, "film_siren_dgrad: g_layer_stride_tiles %lld < %lld tiles", dp.g_stride, dp.T);
  dp.d_feat = d_feat; dp.d_final_b = d_final_b_acc;
  dp.res_save_mask = res_save_mask; dp.res_add_mask = res_add_mask; dp.res_scratch = static_cast<float*>(res_scratch);
  return dgrad_launch(dp, as_stream(stream));
}

int cng_film_siren_wgrad(const void* dz_dump, const void* x_dump, long long x_layer_stride_tiles, const void* feat_dump, long long P, int L,
                         int x_is_fp16, float* const* d_w_acc_host, float* colsum_acc, cng_stream_t stream) {
  using namespace cng;
  using namespace cng::bwdtc;
  CNG_REQUIRE(P >= 0 && L >= 1 && L <= 16, CNG_ERR_INVALID_ARGUMENT, "film_siren_wgrad: bad shape");
  if (P == 0) return CNG_OK;
  CNG_REQUIRE(dz_dump && feat_dump && (L == 1 || x_dump) && d_w_acc_host && colsum_acc, CNG_ERR_INVALID_ARGUMENT, "film_siren_wgrad: NULL pointer");
  CNG_REQUIRE(((reinterpret_cast<uintptr_t>(dz_dump) | reinterpret_cast<uintptr_t>(x_dump) | reinterpret_cast<uintptr_t>(feat_dump)) & 15) == 0,
              CNG_ERR_INVALID_ARGUMENT, "film_siren_wgrad: dumps not 16-byte aligned");
  if (int e = cng_device_check()) return e;
  WgradParams wp{};
  wp.dz = static_cast<const uint8_t*>(dz_dump); wp.x = static_cast<const uint8_t*>(x_dump); wp.feat = static_cast<const uint8_t*>(feat_dump);
  wp.T = (P + kTileM - 1) / kTileM; wp.L = L; wp.x_half = x_is_fp16 ? 1 : 0; wp.colsum = colsum_acc;
  wp.x_stride = x_layer_stride_tiles > 0 ? x_layer_stride_tiles : wp.T;
  CNG_REQUIRE(wp.x_stride >= wp.T, CNG_ERR_INVALID_ARGUMENT, "film_siren_wgrad: x_layer_stride_tiles %lld < %lld tiles", wp.x_stride, wp.T);
  for (int l = 0; l < L; ++l) {
    CNG_REQUIRE(d_w_acc_host[l] && (reinterpret_cast<uintptr_t>(d_w_acc_host[l]) & 15) == 0, CNG_ERR_INVALID_ARGUMENT,
                "film_siren_wgrad: d_w_acc[%d] NULL or not 16-byte aligned", l);
    wp.dW[l] = d_w_acc_host[l];
  }
  return wgrad_launch(wp, as_stream(stream));
}

int cng_film_siren_head_wgrad(const float* d_out, const float* out, int sigmoid_rgb, const void* x_last_tiles, long long P, int x_is_fp16,
                              float* d_final_w_acc, cng_stream_t stream) {
  using namespace cng;
  using namespace cng::bwdtc;
  CNG_REQUIRE(P >= 0, CNG_ERR_INVALID_ARGUMENT, "film_siren_head_wgrad: P=%lld", P);
  if (P == 0) return CNG_OK;
  CNG_REQUIRE(d_out && x_last_tiles && d_final_w_acc && (!sigmoid_rgb || out), CNG_ERR_INVALID_ARGUMENT, "film_siren_head_wgrad: NULL pointer");
  CNG_REQUIRE(((reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(x_last_tiles)) & 15) == 0,
              CNG_ERR_INVALID_ARGUMENT, "film_siren_head_wgrad: buffers not 16-byte aligned");
  if (int e = cng_device_check()) return e;
  const long long T = (P + kTileM - 1) / kTileM;
  const unsigned grid = static_cast<unsigned>(min(static_cast<long long>(3 * sm_count()), T));     // latency-bound streaming: 3 blocks per SM
  if (x_is_fp16) head_wgrad_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(d_out, out, sigmoid_rgb, static_cast<const uint8_t*>(x_last_tiles), P, T, d_final_w_acc);
  else head_wgrad_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(d_out, out, sigmoid_rgb, static_cast<const uint8_t*>(x_last_tiles), P, T, d_final_w_acc);
  return check_launch("cng_film_siren_head_wgrad");
}

int cng_film_siren_bwd(const float* feat, const float* d_out, long long P, int C, int HID, int L, const float* const* layer_w_host,
                       const float* const* layer_b_host, const float* freq, const float* phase, const float* final_w, const float* final_b,
                       int sigmoid_rgb, unsigned res_save_mask, unsigned res_add_mask, void* workspace, size_t workspace_bytes,
                       void* res_scratch, size_t res_scratch_bytes, float* d_feat, float* const* d_w_acc_host, float* colsum_acc,
                       float* d_final_w_acc, float* d_final_b_acc, cng_stream_t stream) {
  using namespace cng;
  using namespace cng::bwdtc;
  CNG_REQUIRE(P >= 0 && C == 32 && HID == kHID && L >= 1 && L <= 16, CNG_ERR_UNSUPPORTED, "film_siren_bwd: needs C=32, HID=256, L<=16 (got %d, %d, %d)", C, HID, L);
  if (P == 0) return CNG_OK;
  CNG_REQUIRE(P < (1LL << 31), CNG_ERR_UNSUPPORTED, "film_siren_bwd: chunk of %lld points (>= 2^31)", P);
  CNG_REQUIRE(feat && d_out && layer_w_host && layer_b_host && freq && phase && final_w && final_b && d_feat && d_w_acc_host && colsum_acc &&
                  d_final_w_acc && d_final_b_acc && workspace,
              CNG_ERR_INVALID_ARGUMENT, "film_siren_bwd: NULL pointer");
  for (int l = 0; l < L; ++l) CNG_REQUIRE(layer_w_host[l] && layer_b_host[l] && d_w_acc_host[l], CNG_ERR_INVALID_ARGUMENT, "film_siren_bwd: NULL layer %d", l);
  const Layout lay = layout(P, L);
  CNG_REQUIRE(workspace_bytes >= lay.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, CNG_ERR_WORKSPACE,
              "film_siren_bwd: workspace %zu < %zu bytes (or not 256-byte aligned)", workspace_bytes, lay.total);
  const bool res = (res_save_mask | res_add_mask) != 0;
  CNG_REQUIRE(!res || (res_scratch && res_scratch_bytes >= cng_film_siren_res_scratch_bytes()), CNG_ERR_WORKSPACE,
              "film_siren_bwd: residual scratch %zu < %zu bytes", res_scratch_bytes, cng_film_siren_res_scratch_bytes());
  if (int e = cng_device_check()) return e;
  cudaStream_t st = as_stream(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  float* out_tmp = reinterpret_cast<float*>(ws + lay.out);
  const long long T = (P + kTileM - 1) / kTileM;
  // 1. recompute with dumps (fp16 operands; x tile images, g in epilogue order, layer-0 operand blocks)
  if (int e = film_siren_tc_train_launch(feat, P, L, layer_w_host, layer_b_host, freq, phase, final_w, final_b, sigmoid_rgb, 1, ws + lay.fold,
                                         film_siren_tc_workspace(1, L), out_tmp, ws + lay.xs, ws + lay.gs, ws + lay.feat, st, res_save_mask,
                                         res_add_mask, static_cast<float*>(res_scratch)))
    return e;
  // 2. W^T operand images
  if (int e = cng_film_siren_wt_images(layer_w_host, freq, final_w, C, HID, L, ws + lay.wt, stream)) return e;
  // 3. dgrad chain: d_feat, d_final_b, dz tile images
  if (int e = cng_film_siren_dgrad(d_out, out_tmp, sigmoid_rgb, P, L, ws + lay.wt, ws + lay.gs, 0, ws + lay.dzs, d_feat, d_final_b_acc, res_save_mask,
                                   res_add_mask, res_scratch, res_scratch_bytes, stream))
    return e;
  // 4. weight gradients + column sums
  if (int e = cng_film_siren_wgrad(ws + lay.dzs, ws + lay.xs, 0, ws + lay.feat, P, L, 1, d_w_acc_host, colsum_acc, stream)) return e;
  // 5. head weights
  return cng_film_siren_head_wgrad(d_out, out_tmp, sigmoid_rgb, ws + lay.xs + static_cast<size_t>(L - 1) * T * kTileImageBytes, P, 1, d_final_w_acc, stream);
}

}  // extern "C"
