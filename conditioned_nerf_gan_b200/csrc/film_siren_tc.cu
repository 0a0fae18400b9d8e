// K2 (throughput mode): FiLM-SIREN MLP on the 5th-gen tensor cores (tcgen05 + TMEM), all layers
// fused, activations never leave the SM.
//
// Replaces, for the FG SIREN family (generators/siren.py:491-580, :583-668, :744-827, :983-1065):
//   per layer  x = sin(freq * (x W^T + b) + phase)      FiLMLayer.forward :146-160
//   head       rgb_sigma = [sigmoid](x Wf^T + bf)        :579 / :1064, _sigmoid_rgb :1227-1234
//
// Math as executed here
//   * FiLM is folded into the operands once per forward (film_fold_kernel): W'_{b,l} =
//     diag(freq_{b,l}) W_l rounded to bf16, shift_{b,l} = freq*b + phase kept in fp32, so the
//     epilogue is sin(acc + shift) on the fp32 TMEM accumulator (the pre-activation is never
//     rounded to bf16 -- SURVEY.md section 7 "precision of the MLP").
//   * layer 0 (K = C = 32) runs as split-bf16: A = [x_hi | x_lo], B = [W'_hi | W'_hi] plus
//     A = [x_hi], B = [W'_lo]  (hi*hi + lo*hi + hi*lo, ~2^-16 relative), because its rounding
//     error is amplified by every later layer.
//   * hidden layers: bf16 x bf16 -> fp32, K = 256 as 4 K-blocks of 64.
//   * head: one N=16 MMA group (rows 4..15 of the operand are zero), bias + sigmoid in the epilogue.
//
// Kernel organisation (one persistent CTA per SM, 576 threads, cta_group::1)
//   warps 0-7   epilogue of tile slot 0     warps 8-15  epilogue of tile slot 1
//               (warp w works on TMEM lanes 32*(w%4).. and on accumulator columns 128*((w>>2)&1)..)
//   warp  16    MMA issuer (one elected lane) + TMEM allocator
//   warp  17    weight producer: cp.async.bulk (TMA unit, SASS UBLKCP) global -> smem ring,
//               completion on mbarriers (complete_tx)
//   The FiLM shift row of the layer in flight (1 KB per tile slot) lives in shared memory and is added to the
//   accumulator in registers (LDS.128, warp-broadcast), so the epilogue per element is FADD + FMUL(1/2pi) +
//   MUFU.SIN + half a pack; the next row is published between two named barriers off the critical chain.
//   Two 128-point tiles are in flight per CTA ("ping-pong"): while the tensor pipe runs layer l
//   of one tile, the epilogue warps of the other tile run sin() on their accumulator and write
//   the next layer's bf16 A operand back into shared memory (128B-swizzled, K-major), so MUFU and
//   tensor pipe overlap.  TMEM: 2 accumulators x 256 fp32 columns = all 512 columns.
//   Shared memory: 2 x 64 KB A tiles + 3 x 32 KB weight ring + barriers = 224.1 KB.
//   Weight K-blocks are stored in global memory as ready-made shared-memory images (already
//   swizzled) so a block is one contiguous 32 KB bulk copy; no tensor map is needed.
//
// Roofline: 2*(32*256*3 + 7*256^2 + 256*16) flop per point issued (algorithmic 935 936), 2048
// sin per point.  Per 128-point tile-layer: 16 MMAs (128x256x16) = 2048 tensor cycles and 32768
// sin = 2048 MUFU cycles at 16/clk/SM: the two pipes are balanced by construction.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "cng_common.cuh"

#include "film_siren_tc_common.cuh"
#include "trilinear.cuh"

#ifndef CNG_TC_EPI_WARPS
#define CNG_TC_EPI_WARPS 8
#endif
#ifndef CNG_TC_EPI_PIPELINE
#define CNG_TC_EPI_PIPELINE 1
#endif
#ifndef CNG_TC_EPI_GROUPED
#define CNG_TC_EPI_GROUPED 0
#endif

namespace cng {

constexpr int kRing = 3;
constexpr int kDefaultCtaGroup = 1;   // measured: the pair kernel pays ~1300 cycles of cross-CTA hand-off per layer
constexpr int kDefaultTrainPolyOneIn = 4;   // training-mode epilogue: one (sin, cos) pair in four on the FMA pipe (measured: 1.93 -> 1.80 ms per 1M points, profiles/r2_bwd_kernels.txt)
constexpr int kDefaultPolyOneIn = 0;   // every sine on the MUFU unit.  One in 8 on the FMA pipe (CNG_TC_POLY=8) won 2 % in round 1 (2.77 -> 2.70 ms per launch); measured again on the round-2 kernel it loses 1-3 % for every class (TALLSIREN_FG 2.72 vs 2.80 ms, SHORTSIREN_FG fp16 1.32 vs 1.36 ms; profiles/r2l_k2_poly_sweep.txt)
constexpr int kEpiWarpsPerSlot = CNG_TC_EPI_WARPS;   // 4, 8 or 12 (build-time knob, see build.py)
static_assert(kEpiWarpsPerSlot == 4 || kEpiWarpsPerSlot == 8 || kEpiWarpsPerSlot == 12, "CNG_TC_EPI_WARPS: 4, 8 or 12");
constexpr int kEpiGroups = kEpiWarpsPerSlot / 4;           // warps per TMEM lane quarter = column groups of the accumulator
constexpr bool kEpiEven = (8 % kEpiGroups) == 0;           // every group owns the same number of 32-column blocks
constexpr int kBlocksPerWarp = (8 + kEpiGroups - 1) / kEpiGroups;   // 32-column accumulator blocks per epilogue warp (the most any group owns)
constexpr int kMmaWarp = 2 * kEpiWarpsPerSlot;
constexpr int kProducerWarp = kMmaWarp + 1;
constexpr int kNumThreads = 32 * (kProducerWarp + 1);
constexpr uint32_t kSmemA = 0;
constexpr uint32_t kSmemW = 2 * kATileBytes;                         // 131072
constexpr uint32_t kSmemBar = kSmemW + kRing * kChunkBytes;          // 229376
constexpr uint32_t kSmemShift = kSmemBar + 128;                      // FiLM shift row of the layer each slot is working on: [2][256] fp32
constexpr uint32_t kSmemTotal = kSmemShift + 2 * kHID * 4;           // 231552 <= 232448

// The two measured-slower alternative organisations (film_siren_tc2.cu: CTA pairs / cta_group::2; film_siren_tc3.cu: one tile per
// CTA, layer-pipelined) are compiled only with CNG_BUILD_EXPERIMENTAL=1 (build.py defines CNG_WITH_EXPERIMENTAL_K2).
#ifdef CNG_WITH_EXPERIMENTAL_K2
int film_siren_tc2_launch(TcParams p, cudaStream_t stream);   // film_siren_tc2.cu
int film_siren_tc3_launch(TcParams p, int poly, cudaStream_t stream);   // film_siren_tc3.cu
#endif
// 1: the two-tile ping-pong kernel below, 8 epilogue warps bound to each tile slot; 2: the same with all 16 epilogue warps
// shared between the slots; 3 (experimental builds only): the layer-pipelined kernel (film_siren_tc3.cu).  All give bit-identical results.
constexpr int kDefaultKernelVersion = 1;   // measured (profiles/r1g_k2_versions.txt): 1 and 2 within 1-2 % on TALLSIREN_FG, 1 ahead for fewer layers, 3 behind by 10 %

// fused-gather mode of film_siren_tc_launch: the features are looked up by the kernel itself
struct GatherSource {
  const float* vol_ndhwc;
  long long vol_item_stride;     // floats; 0 = shared volume
  int D, H, W;
  const float* points;           // [B, N, 3]
};

static int g_dump_bits_override = 0;   // debug hook / tests: cng_internal_set_g_dump_bits
int film_siren_g_dump_bits() {
  if (g_dump_bits_override) return g_dump_bits_override;
  static const int bits = [] {
    const char* e = getenv("CNG_G_DUMP_BITS");
    return (e && atoi(e) == 8) ? 8 : 16;
  }();
  return bits;
}
static long long* g_tc_trace = nullptr;   // debug hook, see cng_internal_set_tc_trace
static int g_tc_version = 0;              // 0: CNG_TC_V / default; 1 or 3: forced (cng_internal_set_tc_version, tests and A/B tools)

// ---- fold kernel ---------------------------------------------------------------------------------
struct FoldParams {
  const float* w[16];
  const float* b[16];
  const float* freq;      // [B, L*HID]
  const float* phase;     // [B, L*HID]
  const float* final_w;   // [4, HID]
  int B, L;
  uint8_t* images;        // [B][item_image_bytes]
  float* shift;           // [B][L][HID]
};

// one thread per (item, layer, row n, 8 consecutive k)
template <bool kHalf>
__global__ void __launch_bounds__(256) film_fold_kernel(FoldParams p) {
  const int L = p.L;
  const long long per_item = 256LL * 4 /*layer0: 32/8*/ + static_cast<long long>(L - 1) * 256 * 32 + 16 * 32 /*head*/;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= per_item * p.B) return;
  const int item = static_cast<int>(gid / per_item);
  long long e = gid - static_cast<long long>(item) * per_item;
  uint8_t* img = p.images + static_cast<size_t>(item) * item_image_bytes(L);
  if (e < 256 * 4) {                                   // ---- layer 0: split bf16
    const int n = static_cast<int>(e >> 2), k0 = static_cast<int>(e & 3) * 8;
    const float fq = __ldg(p.freq + (static_cast<size_t>(item) * L + 0) * kHID + n);
    uint16_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = fq * __ldg(p.w[0] + n * kC0 + k0 + j);
      hi[j] = to16<kHalf>(v);
      lo[j] = to16<kHalf>(v - from16<kHalf>(hi[j]));
    }
    uint8_t* c0 = img + chunk_offset(L, 0, 0);
    uint8_t* c1 = img + chunk_offset(L, 0, 1);
    *reinterpret_cast<uint4*>(c0 + sw128_offset(n, k0)) = *reinterpret_cast<uint4*>(hi);        // pairs with x_hi
    *reinterpret_cast<uint4*>(c0 + sw128_offset(n, 32 + k0)) = *reinterpret_cast<uint4*>(hi);   // pairs with x_lo
    *reinterpret_cast<uint4*>(c1 + sw128_offset(n, k0)) = *reinterpret_cast<uint4*>(lo);        // pairs with x_hi
    *reinterpret_cast<uint4*>(c1 + sw128_offset(n, 32 + k0)) = make_uint4(0, 0, 0, 0);
    if (k0 == 0) {
      const float bias = __ldg(p.b[0] + n), ph = __ldg(p.phase + (static_cast<size_t>(item) * L + 0) * kHID + n);
      p.shift[(static_cast<size_t>(item) * L + 0) * kHID + n] = __fadd_rn(__fmul_rn(fq, bias), ph);
    }
    return;
  }
  e -= 256 * 4;
  if (e < static_cast<long long>(L - 1) * 256 * 32) {  // ---- hidden layers
    const int l = 1 + static_cast<int>(e / (256 * 32));
    const int r = static_cast<int>(e % (256 * 32));
    const int n = r >> 5, k0 = (r & 31) * 8;
    const float fq = __ldg(p.freq + (static_cast<size_t>(item) * L + l) * kHID + n);
    uint16_t hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hi[j] = to16<kHalf>(fq * __ldg(p.w[l] + n * kHID + k0 + j));
    uint8_t* c = img + chunk_offset(L, l, k0 >> 6);
    *reinterpret_cast<uint4*>(c + sw128_offset(n, k0 & 63)) = *reinterpret_cast<uint4*>(hi);
    if (k0 == 0) {
      const float bias = __ldg(p.b[l] + n), ph = __ldg(p.phase + (static_cast<size_t>(item) * L + l) * kHID + n);
      p.shift[(static_cast<size_t>(item) * L + l) * kHID + n] = __fadd_rn(__fmul_rn(fq, bias), ph);
    }
    return;
  }
  e -= static_cast<long long>(L - 1) * 256 * 32;
  {                                                    // ---- head: [16 n][256 k], rows >= 4 zero
    const int n = static_cast<int>(e >> 5), k0 = static_cast<int>(e & 31) * 8;
    uint16_t hi[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) hi[j] = to16<kHalf>(n < 4 ? __ldg(p.final_w + n * kHID + k0 + j) : 0.f);
    uint8_t* c = img + chunk_offset(L, L, 0) + (k0 >> 6) * 2048;
    *reinterpret_cast<uint4*>(c + sw128_offset(n, k0 & 63)) = *reinterpret_cast<uint4*>(hi);
  }
}

// kTrain: additionally stream x_{l+1} and g_l = freq*cos(u_l) of every layer to HBM (p.dump_x / p.dump_g) for the backward
// kShared: all 16 epilogue warps work on whichever tile's accumulator is complete (slot 0, then slot 1, then slot 0 ...)
// instead of 8 warps bound to each slot.  A tile's epilogue then takes about as long as the other tile's MMAs and the
// two strictly alternate: the tensor pipe runs back to back ("what the K2 numbers taught").
// kRes: residual blocks (res_save_mask / res_add_mask of TcParams).  A separate instantiation: merely having the branch in the
// epilogue costs the plain networks 4-5 % (same-box A/B, 2.95 vs 2.81 ms).
template <int kPolyOneIn, bool kHalf, bool kTrain = false, bool kShared = false, bool kRes = false, bool kGather = false, bool kG8 = false>
__global__ void __launch_bounds__(kNumThreads, 1) film_siren_tc_kernel(TcParams p) {
  static_assert(!(kShared && kRes), "residual blocks run on the slot-bound epilogue");
  static_assert(!kGather || (!kTrain && !kShared && !kRes), "the fused gather is an inference path of the slot-bound kernel");
  static_assert(!(kShared && kTrain), "the shared epilogue is an inference path");
  // (built with CNG_TC_EPI_WARPS=4 the shared mode is not instantiated: shared_kernel() below returns nullptr)
  constexpr int kRingN = kRing;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  const uint32_t bar0 = s_base + kSmemBar;
  // barriers: w_full[3] @0, w_empty[3] @24, act_ready[2] @48, acc_full[2] @64, tmem ptr @96
  auto w_full = [&](int s) { return bar0 + 8u * s; };
  auto w_empty = [&](int s) { return bar0 + 24u + 8u * s; };
  auto act_ready = [&](int x) { return bar0 + 48u + 8u * x; };
  auto acc_full = [&](int x) { return bar0 + 64u + 8u * x; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 96);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRingN; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int x = 0; x < 2; ++x) { mbar_init(act_ready(x), kShared ? 2 * kEpiWarpsPerSlot : 32 * kEpiWarpsPerSlot); mbar_init(acc_full(x), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_base + kSmemBar + 96), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const long long G = gridDim.x;
  const long long first = blockIdx.x;

  if (warp == kProducerWarp) {
    // =========================== weight producer ===========================
    {
      const bool elected = elect_one();
      int slot = 0;
      uint32_t phase = 0;
      for (long long t0 = first; t0 < p.total_tiles; t0 += 2 * G) {
        const int nx = (t0 + G < p.total_tiles) ? 2 : 1;
        const int item0 = static_cast<int>(t0 / p.tiles_per_item);
        const int item1 = nx == 2 ? static_cast<int>((t0 + G) / p.tiles_per_item) : 0;
        for (int l = 0; l <= L; ++l) {
          const int nchunks = (l == 0) ? 2 : (l < L ? 4 : 1);
          const uint32_t bytes = (l < L) ? kChunkBytes : kHeadBytes;
          for (int x = 0; x < nx; ++x) {
            const uint8_t* img = p.images + static_cast<size_t>(x == 0 ? item0 : item1) * item_image_bytes(L);
            for (int c = 0; c < nchunks; ++c) {
              mbar_wait(w_empty(slot), phase ^ 1);
              if (elected) {
                mbar_arrive_expect_tx(w_full(slot), bytes);
                bulk_g2s_hint(s_base + kSmemW + slot * kChunkBytes, img + chunk_offset(L, l, c), bytes, w_full(slot), kL2EvictLast);
              }
              __syncwarp();
              if (++slot == kRingN) { slot = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer ===========================
    // The whole warp runs the loop (warp-uniform control flow keeps descriptors and barrier addresses in
    // uniform registers); one elected lane issues the tcgen05 instructions.  Measured with the lane-0-only
    // version: ~150 cycles per MMA issue + ~130 per commit on the issuing thread, i.e. issue-bound.
    {
      const bool elected = elect_one();
      int slot = 0;
      uint32_t phase = 0, act_phase = 0;           // act_phase: bit x = parity of act_ready(x)
      constexpr uint32_t idesc_main = make_idesc(128, 256, kHalf);
      constexpr uint32_t idesc_head = make_idesc(128, 16, kHalf);
      int iter = 0;
      for (long long t0 = first; t0 < p.total_tiles; t0 += 2 * G, ++iter) {
        const int nx = (t0 + G < p.total_tiles) ? 2 : 1;
        for (int l = 0; l <= L; ++l) {
          const int nchunks = (l == 0) ? 2 : (l < L ? 4 : 1);
          for (int x = 0; x < nx; ++x) {
            mbar_wait_lean(act_ready(x), (act_phase >> x) & 1u);
            act_phase ^= 1u << x;
            tc_fence_after();
            if (elected) trace_event(p.trace, iter, l, x, 0);
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(x) * kHID;
            const uint64_t a_desc0 = make_desc(s_base + kSmemA + x * kATileBytes);
            for (int c = 0; c < nchunks; ++c) {
              mbar_wait_lean(w_full(slot), phase);
              tc_fence_after();
              const uint64_t b_desc = make_desc(s_base + kSmemW + slot * kChunkBytes);
              if (elected) {
                if (l < L) {
                  // layer 0: chunk 0 = 4 k-steps over [x_hi|x_lo], chunk 1 = 2 k-steps over [x_hi]
                  const uint64_t a_desc = a_desc0 + (l == 0 ? 0 : c * (kABlockBytes >> 4));
                  tc_mma_bf16(d_tmem, a_desc, b_desc, idesc_main, c ? 1u : 0u);        // first k-step of the layer overwrites
                  tc_mma_bf16(d_tmem, a_desc + 2, b_desc + 2, idesc_main, 1u);         // +32 bytes = one K step of 16 bf16
                  if (!(l == 0 && c == 1)) {
                    tc_mma_bf16(d_tmem, a_desc + 4, b_desc + 4, idesc_main, 1u);
                    tc_mma_bf16(d_tmem, a_desc + 6, b_desc + 6, idesc_main, 1u);
                  }
                } else {
#pragma unroll
                  for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                      tc_mma_bf16(d_tmem, a_desc0 + kb * (kABlockBytes >> 4) + 2 * ks, b_desc + kb * (2048 >> 4) + 2 * ks, idesc_head,
                                  (kb | ks) ? 1u : 0u);
                }
                tc_commit(w_empty(slot));          // slot is free once these MMAs have read it
              }
              __syncwarp();
              if (++slot == kRingN) { slot = 0; phase ^= 1; }
            }
            if (elected) {
              tc_commit(acc_full(x));              // accumulator of tile x complete
              trace_event(p.trace, iter, l, x, 1);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if constexpr (kShared) {
    // =========================== epilogue warps, shared between the two tile slots ===========================
    const int q = warp & 3;                       // TMEM lane quarter == warp_id % 4
    const int cg = warp >> 2;                     // column group: accumulator blocks [2*cg, 2*cg + 2) of 32 columns
    const int row = q * 32 + lane;
    const int tid = threadIdx.x;                  // 0..511
    uint32_t acc_phase = 0;                       // bit x = parity of acc_full(x)
    int iter = 0;
    const bool tracer = warp == 0 && lane == 0;
    float* rows = reinterpret_cast<float*>(smem + kSmemShift);          // [2 slots][256]: shift row of the layer in flight
    float row_next[2] = {0.f, 0.f};                                     // threads 0..255: element tid of slot 0 / 1's next row
    auto signal = [&](uint32_t bar) {
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    auto publish_row = [&](int x) {
      named_bar_sync(1, 512);                     // every epilogue warp is done reading slot x's old row
      if (tid < kHID) rows[x * kHID + tid] = row_next[x];
      named_bar_sync(1, 512);
    };
    for (long long t0 = first; t0 < p.total_tiles; t0 += 2 * G, ++iter) {
      const int nx = (t0 + G < p.total_tiles) ? 2 : 1;
      TileInfo ti[2];
      const float* shift_item[2];
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        ti[x] = tile_info(p, x < nx ? t0 + x * G : t0);
        shift_item[x] = p.shift + static_cast<size_t>(ti[x].item) * L * kHID;
      }
      // ---- per slot: layer-0 shift row, features -> A block 0 as [x_hi(32) | x_lo(32)] ----
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (x >= nx) continue;
        if (iter == 0 && tid < kHID) row_next[x] = __ldg(shift_item[x] + tid);
        publish_row(x);
        const float4* f = reinterpret_cast<const float4*>(p.feat + (static_cast<size_t>(ti[x].item) * p.N + ti[x].n0) * kC0);
        const uint32_t a_base = kSmemA + x * kATileBytes;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int idx = tid + 512 * i, r = idx >> 3, c4 = idx & 7;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < ti[x].rows) v = __ldg(f + idx);
          uint2 hi, lo;
          hi.x = pack2<kHalf>(v.x, v.y); hi.y = pack2<kHalf>(v.z, v.w);
          lo.x = pack2<kHalf>(v.x - from16<kHalf>(to16<kHalf>(v.x)), v.y - from16<kHalf>(to16<kHalf>(v.y)));
          lo.y = pack2<kHalf>(v.z - from16<kHalf>(to16<kHalf>(v.z)), v.w - from16<kHalf>(to16<kHalf>(v.w)));
          *reinterpret_cast<uint2*>(smem + a_base + sw128_offset(r, 4 * c4)) = hi;
          *reinterpret_cast<uint2*>(smem + a_base + sw128_offset(r, 32 + 4 * c4)) = lo;
        }
        signal(act_ready(x));
      }
      for (int l = 0; l < L; ++l) {
        const bool more = l + 1 < L;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          if (x >= nx) continue;
          if (tid < kHID) {                        // next row of this slot: layer l+1 of this tile, or layer 0 of its next tile
            const long long tn = t0 + x * G + 2 * G;
            const float* src = more ? shift_item[x] + (l + 1) * kHID
                                    : p.shift + static_cast<size_t>(tn < p.total_tiles ? tn / p.tiles_per_item : ti[x].item) * L * kHID;
            row_next[x] = __ldg(src + tid);
          }
          const uint32_t a_base = kSmemA + x * kATileBytes;
          const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(x) * kHID;
          const float* row_x = rows + x * kHID;
          mbar_wait(acc_full(x), (acc_phase >> x) & 1u);
          acc_phase ^= 1u << x;
          tc_fence_after();
          if (tracer) trace_event(p.trace, iter, l, x, 2);
          uint32_t va[32], vb[32];
          CNG_TMEM_LD_32(t_lane + (2 * cg) * 32, va);
          tmem_ld_wait();
          CNG_TMEM_LD_32(t_lane + (2 * cg + 1) * 32, vb);
          auto finish = [&](uint32_t (&v)[32], int cc) {
            const float4* rs = reinterpret_cast<const float4*>(row_x + cc * 32);
            uint8_t* blk = smem + a_base + (cc >> 1) * kABlockBytes + row * 128;
#pragma unroll
            for (int i = 0; i < 4; ++i) {                       // one 16-byte chunk (8 columns) at a time
              const float4 s0 = rs[2 * i], s1 = rs[2 * i + 1];
              const float sh[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 8; j += 2)
                o[j / 2] = pack2<kHalf>(film_sin<kPolyOneIn>(__uint_as_float(v[8 * i + j]) + sh[j], 8 * i + j),
                                        film_sin<kPolyOneIn>(__uint_as_float(v[8 * i + j + 1]) + sh[j + 1], 8 * i + j + 1));
              const int chunk = ((cc & 1) * 4 + i) ^ (row & 7);
              *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          };
          finish(va, 2 * cg);
          tmem_ld_wait();
          finish(vb, 2 * cg + 1);
          if (tracer) trace_event(p.trace, iter, l, x, 3);
          signal(act_ready(x));
          if (more) publish_row(x);        // the next tile's layer-0 row is published at the top of the tile loop
        }
      }
      // ---- head: 4 accumulator columns -> bias, sigmoid(rgb), store (column group 0 holds them) ----
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (x >= nx) continue;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(x) * kHID;
        mbar_wait(acc_full(x), (acc_phase >> x) & 1u);
        acc_phase ^= 1u << x;
        tc_fence_after();
        if (cg == 0) {
          uint32_t r0, r1, r2, r3;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                       : "r"(t_lane)
                       : "memory");
          tmem_ld_wait();
          float4 o;
          const float4 fb = __ldg(reinterpret_cast<const float4*>(p.final_b));
          o.x = __uint_as_float(r0) + fb.x;
          o.y = __uint_as_float(r1) + fb.y;
          o.z = __uint_as_float(r2) + fb.z;
          o.w = __uint_as_float(r3) + fb.w;
          if (p.sigmoid_rgb) {
            o.x = 1.f / (1.f + __expf(-o.x));
            o.y = 1.f / (1.f + __expf(-o.y));
            o.z = 1.f / (1.f + __expf(-o.z));
          }
          if (row < ti[x].rows) reinterpret_cast<float4*>(p.out)[static_cast<size_t>(ti[x].item) * p.N + ti[x].n0 + row] = o;
        }
        tc_fence_before();   // orders the TMEM reads above before the next tile's first MMA (via act_ready)
      }
    }
  } else {
    // =========================== epilogue warps (slot x = warp / 8) ===========================
    const int x = warp / kEpiWarpsPerSlot;
    const int q = warp & 3;                       // TMEM lane quarter == warp_id % 4
    const int half = (warp % kEpiWarpsPerSlot) >> 2;   // column group: accumulator blocks [cc0, cc1) of 32 columns (4 + 4, or 2 + 3 + 3 with 12 warps)
    constexpr int kB = kBlocksPerWarp;
    const int cc0 = (8 * half) / kEpiGroups, cc1 = (8 * (half + 1)) / kEpiGroups;
    const int row = q * 32 + lane;
    const uint32_t a_base = kSmemA + x * kATileBytes;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(x) * kHID;
    uint32_t acc_phase = 0;
    int iter = 0;
    const bool tracer = (warp % kEpiWarpsPerSlot) == 0 && lane == 0;
    // FiLM shift: the row of the layer in flight lives in shared memory (1 KB per slot) and is added to the accumulator
    // in registers (LDS.128, warp-broadcast).  The next row is fetched into a register at the start of a layer's
    // epilogue and published after it, between two barriers of the slot's warps, while the tensor pipe runs the next
    // layer -- off the critical chain.  (Measured on the pipelined kernel: pre-storing the shift in the accumulator
    // with tcgen05.st put an L2-latency load and a TMEM store on every block's chain, ~900 cycles per layer.)
    constexpr int kSlotThreads = 32 * kEpiWarpsPerSlot;
    constexpr int kRowPerThread = (kHID + kSlotThreads - 1) / kSlotThreads;
    const int ts = (warp % kEpiWarpsPerSlot) * 32 + lane;
    float* row_x = reinterpret_cast<float*>(smem + kSmemShift) + x * kHID;
    float row_next[kRowPerThread];
    constexpr bool res_mode = kRes;
    auto publish_row = [&]() {
      named_bar_sync(1 + x, kSlotThreads);            // every warp of the slot is done reading the old row
#pragma unroll
      for (int i = 0; i < kRowPerThread; ++i)
        if (ts + i * kSlotThreads < kHID) row_x[ts + i * kSlotThreads] = row_next[i];
      named_bar_sync(1 + x, kSlotThreads);
    };
    // training mode: the A tile leaves for HBM as one bulk store per layer (issued by one thread of the slot); before the
    // tile is written again that store must have finished READING shared memory
    const bool storer = (warp % kEpiWarpsPerSlot) == 0 && lane == 0;
    auto tile_free = [&]() {
      if (storer) bulk_wait_read_all();
      named_bar_sync(1 + x, kSlotThreads);
    };
    for (long long t = first + x * G; t < p.total_tiles; t += 2 * G, ++iter) {
      const TileInfo ti = tile_info(p, t);
      const float* shift_item = p.shift + static_cast<size_t>(ti.item) * L * kHID;
      // ---- shift row of layer 0 -> shared memory (prefetched during the previous tile's last layer, see below) ----
      if (iter == 0) {
#pragma unroll
        for (int i = 0; i < kRowPerThread; ++i)
          if (ts + i * kSlotThreads < kHID) row_next[i] = __ldg(shift_item + ts + i * kSlotThreads);
      }
      publish_row();
      if constexpr (kTrain) tile_free();
      // ---- features -> A block 0 as [x_hi(32) | x_lo(32)] ----
      {
        const float4* f = kGather ? nullptr : reinterpret_cast<const float4*>(p.feat + (static_cast<size_t>(ti.item) * p.N + ti.n0) * kC0);
        const float* pts = kGather ? p.points + (static_cast<size_t>(ti.item) * p.N + ti.n0) * 3 : nullptr;
        const float4* vol = kGather ? p.vol + static_cast<size_t>(ti.item) * p.vol_item_stride : nullptr;
#pragma unroll
        for (int it = cc0; it < cc1; ++it) {
          const int r = q * 32 + it * 4 + (lane >> 3);
          const int c4 = lane & 7;                               // float4 index within the row: k = 4*c4
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (kGather) {
            if (r < ti.rows) {                                   // 8 lanes x float4 = the 32 channels of one point, 8 corner loads per lane
              const CornerRec rec = corner_record(__ldg(pts + 3 * r), __ldg(pts + 3 * r + 1), __ldg(pts + 3 * r + 2), p.D, p.H, p.W);
              v = gather_c4(vol, p.H, p.W, kC0 / 4, rec, c4);
            }
          } else {
            if (r < ti.rows) v = __ldg(f + r * 8 + c4);
          }
          uint2 hi, lo;
          hi.x = pack2<kHalf>(v.x, v.y); hi.y = pack2<kHalf>(v.z, v.w);
          lo.x = pack2<kHalf>(v.x - from16<kHalf>(to16<kHalf>(v.x)), v.y - from16<kHalf>(to16<kHalf>(v.y)));
          lo.y = pack2<kHalf>(v.z - from16<kHalf>(to16<kHalf>(v.z)), v.w - from16<kHalf>(to16<kHalf>(v.w)));
          *reinterpret_cast<uint2*>(smem + a_base + sw128_offset(r, 4 * c4)) = hi;
          *reinterpret_cast<uint2*>(smem + a_base + sw128_offset(r, 32 + 4 * c4)) = lo;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      if constexpr (kTrain) {                                      // the layer-0 operand block -> HBM (operand of the layer-0 weight gradient)
        named_bar_sync(1 + x, kSlotThreads);
        if (storer) {
          bulk_s2g_hint(p.dump_feat + static_cast<size_t>(t) * kABlockBytes, s_base + a_base, kABlockBytes, kL2EvictFirst);
          bulk_commit();
        }
      }
      mbar_arrive(act_ready(x));
      for (int l = 0; l < L; ++l) {
        const bool more = l + 1 < L;
        {                                                          // next row: this tile's layer l+1, or the next tile's layer 0
          const long long tn = t + 2 * G;
          const float* src = more ? shift_item + (l + 1) * kHID
                                  : p.shift + static_cast<size_t>(tn < p.total_tiles ? tn / p.tiles_per_item : ti.item) * L * kHID;
#pragma unroll
          for (int i = 0; i < kRowPerThread; ++i)
            if (ts + i * kSlotThreads < kHID) row_next[i] = __ldg(src + ts + i * kSlotThreads);
        }
        mbar_wait(acc_full(x), acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (tracer) trace_event(p.trace, iter, l, x, 2);
        if constexpr (kTrain) tile_free();
        auto finish_block = [&](uint32_t (&v)[32], int cc) {
          {
            const float4* rs = reinterpret_cast<const float4*>(row_x + cc * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 sv = rs[i];
              v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + sv.x);
              v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + sv.y);
              v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + sv.z);
              v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + sv.w);
            }
          }
          if constexpr (res_mode) {                                // only the residual SIREN variants' instantiations
            float4* rsd = reinterpret_cast<float4*>(p.res_scratch) + (static_cast<size_t>(blockIdx.x) * 2 + x) * (kHID / 4) * kTileM;
            if ((p.res_add_mask >> l) & 1u) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 r4 = rsd[static_cast<size_t>(cc * 8 + i) * kTileM + row];
                v[4 * i] = __float_as_uint(__uint_as_float(v[4 * i]) + r4.x);
                v[4 * i + 1] = __float_as_uint(__uint_as_float(v[4 * i + 1]) + r4.y);
                v[4 * i + 2] = __float_as_uint(__uint_as_float(v[4 * i + 2]) + r4.z);
                v[4 * i + 3] = __float_as_uint(__uint_as_float(v[4 * i + 3]) + r4.w);
              }
            }
            if ((p.res_save_mask >> l) & 1u) {                     // this layer's output in fp32 (the operand copy below is 16-bit)
#pragma unroll
              for (int i = 0; i < 8; ++i)
                rsd[static_cast<size_t>(cc * 8 + i) * kTileM + row] =
                    make_float4(__sinf(__uint_as_float(v[4 * i])), __sinf(__uint_as_float(v[4 * i + 1])),
                                __sinf(__uint_as_float(v[4 * i + 2])), __sinf(__uint_as_float(v[4 * i + 3])));
            }
          }
          if constexpr (kTrain) {
            // sin -> next layer's operand (shared memory A tile; the whole tile leaves as one bulk store after the layer);
            // cos -> HBM straight from registers in the epilogue's own order, 512 contiguous bytes per warp and store instruction;
            // the dgrad kernel reads it back with the same mapping.  Two formats (kG8; as a run-time branch the pair cost both 10 %):
            //   16: fp16, [cc][q][i][lane] x 16 B -- 512 B per point and layer;
            //    8: u = round(127 cos) + 128, [cc][q][h][lane] x 16 B (h = columns 16 h .. 16 h + 15 of the block) -- 256 B per point
            //       and layer: the kernel is bound by its HBM writes (1.80 -> 1.66 ms per 1 M points), at the price of a
            //       quantisation error of <= 1/254 per element
            constexpr bool g8 = kG8;
            uint4* gt = reinterpret_cast<uint4*>(p.dump_g + (static_cast<size_t>(l) * p.total_tiles + t) * (g8 ? kTileM * kHID : kTileM * kHID * 2));
            uint8_t* blk = smem + a_base + (cc >> 1) * kABlockBytes + row * 128;
            uint32_t gs[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint32_t xo[4], gq[4];
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                float s0, c0, s1, c1;
                film_sincos<kPolyOneIn>(__uint_as_float(v[8 * i + j]), 8 * i + j, s0, c0);
                film_sincos<kPolyOneIn>(__uint_as_float(v[8 * i + j + 1]), 8 * i + j + 1, s1, c1);
                xo[j / 2] = pack2<kHalf>(s0, s1);                  // next layer's tensor-core operand == the x dump
                // 8 bits: the low byte of (127 c + 1.5 * 2^23 + 128) = round(127 c) + 128
                gq[j / 2] = g8 ? __byte_perm(__float_as_uint(fmaf(c0, 127.f, 12583040.f)), __float_as_uint(fmaf(c1, 127.f, 12583040.f)), 0x0040)
                               : pack2<true>(c0, c1);
              }
              const int chunk = ((cc & 1) * 4 + i) ^ (row & 7);
              *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(xo[0], xo[1], xo[2], xo[3]);
              if constexpr (g8) {
                gs[2 * i] = __byte_perm(gq[0], gq[1], 0x5410);
                gs[2 * i + 1] = __byte_perm(gq[2], gq[3], 0x5410);
              } else {
                st_global_evict_first(gt + ((cc * 4 + q) * 4 + i) * 32 + lane, make_uint4(gq[0], gq[1], gq[2], gq[3]));
              }
            }
            if constexpr (g8) {
              st_global_evict_first(gt + ((cc * 4 + q) * 2 + 0) * 32 + lane, make_uint4(gs[0], gs[1], gs[2], gs[3]));
              st_global_evict_first(gt + ((cc * 4 + q) * 2 + 1) * 32 + lane, make_uint4(gs[4], gs[5], gs[6], gs[7]));
            }
            return;
          }
          // 32 columns = 64 bytes = 4 x 16-byte chunks of K-block cc/2, logical chunk (cc&1)*4 + i
          uint8_t* blk = smem + a_base + (cc >> 1) * kABlockBytes + row * 128;
#if CNG_TC_EPI_GROUPED
#pragma unroll
          for (int i = 0; i < 4; ++i) {                 // one chunk (8 sines, 4 packs, 1 store) at a time
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 8; j += 2)
              o[j / 2] = pack2<kHalf>(film_sin<kPolyOneIn>(__uint_as_float(v[8 * i + j]), 8 * i + j),
                                      film_sin<kPolyOneIn>(__uint_as_float(v[8 * i + j + 1]), 8 * i + j + 1));
            const int chunk = ((cc & 1) * 4 + i) ^ (row & 7);
            *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(o[0], o[1], o[2], o[3]);
          }
#else
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2)
            o[j / 2] = pack2<kHalf>(film_sin<kPolyOneIn>(__uint_as_float(v[j]), j), film_sin<kPolyOneIn>(__uint_as_float(v[j + 1]), j + 1));
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int chunk = ((cc & 1) * 4 + i) ^ (row & 7);
            *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
#endif
        };
        if constexpr (kEpiWarpsPerSlot == 4 || CNG_TC_EPI_PIPELINE) {
          // keep the TMEM load of block i+1 in flight under the sines of block i (needs 2 x 32 accumulator registers)
          uint32_t va[32], vb[32];
          CNG_TMEM_LD_32(t_lane + cc0 * 32, va);
#pragma unroll
          for (int i = 0; i < kB; i += 2) {
            const int cc = cc0 + i;
            if (kEpiEven || cc < cc1) {
              const bool more1 = kEpiEven ? (i + 1 < kB) : (cc + 1 < cc1), more2 = kEpiEven ? (i + 2 < kB) : (cc + 2 < cc1);
              tmem_ld_wait();
              if (more1) CNG_TMEM_LD_32(t_lane + (cc + 1) * 32, vb);
              finish_block(va, cc);
              if (more1) {
                tmem_ld_wait();
                if (more2) CNG_TMEM_LD_32(t_lane + (cc + 2) * 32, va);
                finish_block(vb, cc + 1);
              }
            }
          }
        } else {
#pragma unroll 1
          for (int cc = cc0; cc < cc1; ++cc) {
            uint32_t v[32];
            CNG_TMEM_LD_32(t_lane + cc * 32, v);
            tmem_ld_wait();
            finish_block(v, cc);
          }
        }
        tc_fence_before();
        fence_proxy_async();
        if (tracer) trace_event(p.trace, iter, l, x, 3);
        if constexpr (kTrain) {                                    // x_{l+1} tile image -> HBM
          named_bar_sync(1 + x, kSlotThreads);
          if (storer) {
            bulk_s2g_hint(p.dump_x + (static_cast<size_t>(l) * p.total_tiles + t) * kATileBytes, s_base + a_base, kATileBytes, kL2EvictFirst);
            bulk_commit();
          }
        }
        mbar_arrive(act_ready(x));
        if (more) publish_row();        // the last layer's prefetch (next tile's layer 0) is published at the top of the tile loop
      }
      // ---- head: 4 accumulator columns -> bias, sigmoid(rgb), store (the column-half-0 warps hold them) ----
      mbar_wait(acc_full(x), acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      if (half == 0) {
        uint32_t r0, r1, r2, r3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                     : "r"(t_lane)
                     : "memory");
        tmem_ld_wait();
        float4 o;
        const float4 fb = __ldg(reinterpret_cast<const float4*>(p.final_b));
        o.x = __uint_as_float(r0) + fb.x;
        o.y = __uint_as_float(r1) + fb.y;
        o.z = __uint_as_float(r2) + fb.z;
        o.w = __uint_as_float(r3) + fb.w;
        if (p.sigmoid_rgb) {
          o.x = 1.f / (1.f + __expf(-o.x));
          o.y = 1.f / (1.f + __expf(-o.y));
          o.z = 1.f / (1.f + __expf(-o.z));
        }
        if (row < ti.rows) reinterpret_cast<float4*>(p.out)[static_cast<size_t>(ti.item) * p.N + ti.n0 + row] = o;
      }
      tc_fence_before();   // orders the TMEM reads above before the next tile's stores / first MMA (via act_ready)
    }
    if constexpr (kTrain) {
      if (storer) bulk_wait_all();      // the last tile images have left shared memory before the CTA exits
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// the shared-epilogue instantiations exist only in the default 8-warps-per-slot build
template <int kPoly, bool kHalf>
static void (*shared_kernel())(TcParams) {
  if constexpr (kEpiWarpsPerSlot == 8) return film_siren_tc_kernel<kPoly, kHalf, false, true>;
  else return nullptr;
}

size_t film_siren_tc_workspace(int B, int L) {
  return static_cast<size_t>(B) * item_image_bytes(L) + static_cast<size_t>(B) * L * kHID * sizeof(float);
}

int film_siren_tc_launch(const float* feat, int B, long long N, int C, int HID, int L, const float* const* w,
                         const float* const* b, const float* freq, const float* phase, const float* final_w,
                         const float* final_b_dev, int sigmoid_rgb, int half_operands, void* workspace, size_t workspace_bytes,
                         float* out, void* dump_x, void* dump_g, cudaStream_t stream, unsigned res_save_mask = 0, unsigned res_add_mask = 0,
                         float* res_scratch = nullptr, void* dump_feat = nullptr, const GatherSource* gather = nullptr) {
  CNG_REQUIRE(HID == kHID && C == kC0, CNG_ERR_UNSUPPORTED, "film_siren_fwd(bf16): needs HID=256, C=32 (got %d, %d)", HID, C);
  CNG_REQUIRE(L >= 1 && L <= 16, CNG_ERR_UNSUPPORTED, "film_siren_fwd(bf16): L=%d", L);
  CNG_REQUIRE(workspace != nullptr && workspace_bytes >= film_siren_tc_workspace(B, L), CNG_ERR_WORKSPACE,
              "film_siren_fwd(bf16): workspace %zu < %zu bytes", workspace_bytes, film_siren_tc_workspace(B, L));
  CNG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127) == 0, CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd(bf16): workspace not 128-byte aligned");
  CNG_REQUIRE((reinterpret_cast<uintptr_t>(feat) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, CNG_ERR_INVALID_ARGUMENT,
              "film_siren_fwd(bf16): feat/out not 16-byte aligned");
  CNG_REQUIRE((feat != nullptr) != (gather != nullptr), CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd(bf16): exactly one of gathered features / gather source");
  CNG_REQUIRE(gather == nullptr || (dump_x == nullptr && (res_save_mask | res_add_mask) == 0), CNG_ERR_UNSUPPORTED,
              "film_siren_fwd_gather: inference of the plain FiLM networks only");
  FoldParams fp{};
  for (int l = 0; l < L; ++l) { fp.w[l] = w[l]; fp.b[l] = b[l]; }
  fp.freq = freq; fp.phase = phase; fp.final_w = final_w; fp.B = B; fp.L = L;
  fp.images = static_cast<uint8_t*>(workspace);
  fp.shift = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + static_cast<size_t>(B) * item_image_bytes(L));
  const long long per_item = 256LL * 4 + static_cast<long long>(L - 1) * 256 * 32 + 16 * 32;
  const long long fold_threads = per_item * B;
  if (half_operands) film_fold_kernel<true><<<static_cast<unsigned>((fold_threads + 255) / 256), 256, 0, stream>>>(fp);
  else film_fold_kernel<false><<<static_cast<unsigned>((fold_threads + 255) / 256), 256, 0, stream>>>(fp);
  if (int e = check_launch("cng_film_siren_fwd(bf16): fold")) return e;

  TcParams p{};
  p.half_operands = half_operands;
  p.dump_x = static_cast<uint8_t*>(dump_x); p.dump_g = static_cast<uint8_t*>(dump_g); p.dump_feat = static_cast<uint8_t*>(dump_feat);
  p.g_bits = film_siren_g_dump_bits();
  CNG_REQUIRE((dump_x == nullptr) == (dump_g == nullptr) && (dump_x == nullptr) == (dump_feat == nullptr), CNG_ERR_INVALID_ARGUMENT,
              "film_siren_fwd_train: the x, g and feature dumps go together");
  CNG_REQUIRE(((reinterpret_cast<uintptr_t>(dump_x) | reinterpret_cast<uintptr_t>(dump_g) | reinterpret_cast<uintptr_t>(dump_feat)) & 15) == 0,
              CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_train: dump buffers not 16-byte aligned");
  p.trace = g_tc_trace;
  p.res_save_mask = res_save_mask; p.res_add_mask = res_add_mask; p.res_scratch = res_scratch;
  CNG_REQUIRE(((res_save_mask | res_add_mask) == 0) || res_scratch != nullptr, CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd: residual masks without a scratch buffer");
  CNG_REQUIRE(((res_save_mask | res_add_mask) >> L) == 0, CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd: residual mask bit beyond layer %d", L - 1);
  p.feat = feat; p.N = N; p.B = B; p.L = L; p.images = fp.images; p.shift = fp.shift;
  if (gather != nullptr) {
    p.vol = reinterpret_cast<const float4*>(gather->vol_ndhwc); p.vol_item_stride = gather->vol_item_stride / 4;
    p.D = gather->D; p.H = gather->H; p.W = gather->W; p.points = gather->points;
  }
  CNG_REQUIRE((reinterpret_cast<uintptr_t>(final_b_dev) & 15) == 0, CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd(bf16): final_b not 16-byte aligned");
  p.final_b = final_b_dev;
  cudaError_t ce = cudaSuccess;
  p.sigmoid_rgb = sigmoid_rgb; p.out = out;
  p.tiles_per_item = (N + kTileM - 1) / kTileM;
  p.total_tiles = p.tiles_per_item * B;
  // CTA-pair kernel (film_siren_tc2.cu, cta_group::2) unless CNG_TC_CG=1 asks for the one-CTA-per-SM kernel below
  static const int cta_group = [] {
    const char* e = getenv("CNG_TC_CG");
    const int v = e ? atoi(e) : kDefaultCtaGroup;
    return (v == 1 || v == 2) ? v : kDefaultCtaGroup;
  }();
#ifdef CNG_WITH_EXPERIMENTAL_K2
  if (cta_group == 2 && sm_count() >= 2 && !half_operands && p.dump_x == nullptr) return film_siren_tc2_launch(p, stream);
#else
  (void)cta_group;
#endif
  // share of the sines evaluated on the FMA pipe instead of the MUFU unit (tuning knob, default from measurement)
  static const int poly = [] {
    const char* e = getenv("CNG_TC_POLY");
    const int v = e ? atoi(e) : kDefaultPolyOneIn;
    return (v == 0 || v == 2 || v == 3 || v == 4 || v == 8) ? v : kDefaultPolyOneIn;
  }();
  const bool train = p.dump_x != nullptr;
  static const int version = [] {
    const char* e = getenv("CNG_TC_V");
    const int v = e ? atoi(e) : kDefaultKernelVersion;
    return (v >= 1 && v <= 3) ? v : kDefaultKernelVersion;
  }();
  const int ver = ((res_save_mask | res_add_mask) || gather) ? 1 : (g_tc_version ? g_tc_version : version);     // residual blocks, fused gather: the slot-bound kernel
#ifdef CNG_WITH_EXPERIMENTAL_K2
  if (ver == 3 && !train && L <= 8) return film_siren_tc3_launch(p, poly, stream);
#endif
  const bool shared = (ver == 2) && !train && kEpiWarpsPerSlot == 8;
  using KernelFn = void (*)(TcParams);
  const int pl = (poly == 0 || poly == 4) ? poly : 8;          // shared mode and fp16 come in these three flavours
  const bool res = (res_save_mask | res_add_mask) != 0;
  // share of the (sin, cos) pairs of the training-mode epilogue evaluated on the FMA pipe (it needs two MUFU ops per element otherwise):
  // one in four (default) or none (CNG_TC_TRAIN_POLY=0); and the format of the cos dump (film_siren_g_dump_bits)
  static const int train_poly = [] {
    const char* e = getenv("CNG_TC_TRAIN_POLY");
    return (e && atoi(e) == 0) ? 0 : kDefaultTrainPolyOneIn;
  }();
  const bool g8 = p.g_bits == 8;
#define CNG_TRAIN_FN(POLY, HALF, RES) (g8 ? film_siren_tc_kernel<POLY, HALF, true, false, RES, false, true> : film_siren_tc_kernel<POLY, HALF, true, false, RES, false, false>)
  const KernelFn train_fn = half_operands ? (train_poly == 0 ? CNG_TRAIN_FN(0, true, false) : CNG_TRAIN_FN(4, true, false))
                                          : (train_poly == 0 ? CNG_TRAIN_FN(0, false, false) : CNG_TRAIN_FN(4, false, false));
  const KernelFn train_res_fn = half_operands ? CNG_TRAIN_FN(0, true, true) : CNG_TRAIN_FN(0, false, true);
#undef CNG_TRAIN_FN
  const bool p0 = poly == 0;                                  // residual / fused-gather instantiations: all sines on the MUFU unit, or one in eight on the FMA pipe
  const KernelFn fn = gather ? (half_operands ? (p0 ? film_siren_tc_kernel<0, true, false, false, false, true> : film_siren_tc_kernel<8, true, false, false, false, true>)
                                              : (p0 ? film_siren_tc_kernel<0, false, false, false, false, true> : film_siren_tc_kernel<8, false, false, false, false, true>))
                      : res ? (train ? train_res_fn
                                   : half_operands ? (p0 ? film_siren_tc_kernel<0, true, false, false, true> : film_siren_tc_kernel<8, true, false, false, true>)
                                                   : (p0 ? film_siren_tc_kernel<0, false, false, false, true> : film_siren_tc_kernel<8, false, false, false, true>))
                      : train ? train_fn
                      : shared ? (half_operands ? (pl == 0 ? shared_kernel<0, true>() : pl == 4 ? shared_kernel<4, true>() : shared_kernel<8, true>())
                                                : (pl == 0 ? shared_kernel<0, false>() : pl == 4 ? shared_kernel<4, false>() : shared_kernel<8, false>()))
                      : half_operands ? (pl == 0 ? film_siren_tc_kernel<0, true> : pl == 4 ? film_siren_tc_kernel<4, true> : film_siren_tc_kernel<8, true>)
                      : poly == 0 ? film_siren_tc_kernel<0, false> : poly == 2 ? film_siren_tc_kernel<2, false>
                      : poly == 3 ? film_siren_tc_kernel<3, false> : poly == 4 ? film_siren_tc_kernel<4, false> : film_siren_tc_kernel<8, false>;
  // function attributes are per device: the opt-in is cached per device ordinal (a process may render on several GPUs)
  static bool attr_set[64][8][12] = {};
  const int variant = res ? 5 + (train ? 2 : half_operands ? 1 : 0) : train ? 2 : (half_operands ? 1 : 0) + (shared ? 3 : 0);
  const int gslot = gather ? (p0 ? 11 : 10) : -1;                     // the fused-gather instantiations: variant rows 0 / 1, columns 10 / 11
  const int pslot = gslot >= 0 ? gslot : train ? 0 : poly;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = -1;
  if (train) dev = -1;                                                // the training-mode instantiations (operand format x poly x dump format) are not cached
  if (dev < 0 || !attr_set[dev][variant][pslot]) {
    ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemTotal));
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "film_siren_fwd(bf16): smem attribute: %s", cudaGetErrorString(ce));
    if (dev >= 0) attr_set[dev][variant][pslot] = true;
  }
  const long long grid = min(static_cast<long long>(sm_count()), p.total_tiles);
  fn<<<static_cast<unsigned>(grid), kNumThreads, kSmemTotal, stream>>>(p);
  return check_launch("cng_film_siren_fwd(bf16)");
}

// the backward's recompute: one item, dumps in the formats of TcParams (film_siren_bwd_tc.cu)
int film_siren_tc_train_launch(const float* feat, long long N, int L, const float* const* w, const float* const* b, const float* freq,
                               const float* phase, const float* final_w, const float* final_b_dev, int sigmoid_rgb, int half_operands,
                               void* workspace, size_t workspace_bytes, float* out, void* dump_x, void* dump_g, void* dump_feat,
                               cudaStream_t stream, unsigned res_save_mask, unsigned res_add_mask, float* res_scratch) {
  return film_siren_tc_launch(feat, 1, N, kC0, kHID, L, w, b, freq, phase, final_w, final_b_dev, sigmoid_rgb, half_operands, workspace,
                              workspace_bytes, out, dump_x, dump_g, stream, res_save_mask, res_add_mask, res_scratch, dump_feat);
}

// defined in film_siren_simt.cu
int film_siren_simt_launch(const float* feat, int B, long long N, int C, int HID, int L, const float* const* w,
                           const float* const* b, const float* freq, const float* phase, const float* final_w,
                           const float* final_b, int sigmoid_rgb, float* out, cudaStream_t stream, unsigned res_save_mask = 0,
                           unsigned res_add_mask = 0);

}  // namespace cng

extern "C" {

// Debug hook (not part of the ABI in include/cng_b200.h): device buffer of 4*9*2*8 int64 that receives the clock64
// timeline of CTA 0 of the next cta_group::1 launches; NULL switches it off.  Used by tools/trace_tc.py.
// format of the cos(u) dump of the NEXT training-mode forward / dgrad calls: 8, 16, or 0 = back to CNG_G_DUMP_BITS / the default (16).
// Switch only between complete backward passes: a dump must be read in the format it was written in.
CNG_API void cng_internal_set_g_dump_bits(int bits) { cng::g_dump_bits_override = (bits == 8 || bits == 16) ? bits : 0; }
CNG_API void cng_internal_set_tc_trace(void* dev_buffer) { cng::g_tc_trace = static_cast<long long*>(dev_buffer); }

// Debug hook: which tcgen05 kernel serves cng_film_siren_fwd -- 1 / 2 the two-tile ping-pong kernel (this file) with slot-bound /
// shared epilogue warps, 3 the layer-pipelined kernel (film_siren_tc3.cu), 0 back to CNG_TC_V / the built-in default.
CNG_API void cng_internal_set_tc_version(int v) { cng::g_tc_version = (v >= 1 && v <= 3) ? v : 0; }

size_t cng_film_siren_workspace_bytes(int B, int C, int HID, int L, int precision) {
  (void)C; (void)HID;
  if ((precision != CNG_PREC_BF16 && precision != CNG_PREC_FP16) || B <= 0 || L <= 0) return 0;
  return cng::film_siren_tc_workspace(B, L);
}

int cng_film_siren_fwd(const float* feat, int B, long long N, int C, int HID, int L, const float* const* layer_w_host,
                       const float* const* layer_b_host, const float* freq, const float* phase, const float* final_w,
                       const float* final_b, int sigmoid_rgb, int precision, void* workspace, size_t workspace_bytes,
                       float* rgb_sigma, cng_stream_t stream) {
  CNG_REQUIRE(B >= 0 && N >= 0 && C >= 1 && HID >= 1 && L >= 1, CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd: bad shape");
  if (B == 0 || N == 0) return CNG_OK;
  CNG_REQUIRE(feat && layer_w_host && layer_b_host && freq && phase && final_w && final_b && rgb_sigma,
              CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd: NULL pointer");
  for (int l = 0; l < L && l < 16; ++l)
    CNG_REQUIRE(layer_w_host[l] && layer_b_host[l], CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd: NULL layer %d", l);
  if (B == 0 || N == 0) return CNG_OK;
  if (int e = cng_device_check()) return e;
  if (precision == CNG_PREC_FP32)
    return cng::film_siren_simt_launch(feat, B, N, C, HID, L, layer_w_host, layer_b_host, freq, phase, final_w, final_b,
                                       sigmoid_rgb, rgb_sigma, cng::as_stream(stream));
  if (precision == CNG_PREC_BF16 || precision == CNG_PREC_FP16)
    return cng::film_siren_tc_launch(feat, B, N, C, HID, L, layer_w_host, layer_b_host, freq, phase, final_w, final_b,
                                     sigmoid_rgb, precision == CNG_PREC_FP16 ? 1 : 0, workspace, workspace_bytes, rgb_sigma,
                                     nullptr, nullptr, cng::as_stream(stream));
  return cng::fail(CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd: unknown precision %d", precision);
}

int cng_film_siren_fwd_gather(const float* vol_ndhwc, long long vol_item_stride, int D, int H, int W, const float* points, int B, long long N,
                              int C, int HID, int L, const float* const* layer_w_host, const float* const* layer_b_host, const float* freq,
                              const float* phase, const float* final_w, const float* final_b, int sigmoid_rgb, int precision,
                              void* workspace, size_t workspace_bytes, float* rgb_sigma, cng_stream_t stream) {
  CNG_REQUIRE(B >= 0 && N >= 0 && C == cng::kC0 && HID == cng::kHID && L >= 1, CNG_ERR_UNSUPPORTED, "film_siren_fwd_gather: needs C=32, HID=256");
  if (B == 0 || N == 0) return CNG_OK;
  CNG_REQUIRE(vol_ndhwc && points && layer_w_host && layer_b_host && freq && phase && final_w && final_b && rgb_sigma, CNG_ERR_INVALID_ARGUMENT,
              "film_siren_fwd_gather: NULL pointer");
  CNG_REQUIRE(D >= 1 && H >= 1 && W >= 1 && (reinterpret_cast<uintptr_t>(vol_ndhwc) & 15) == 0, CNG_ERR_INVALID_ARGUMENT,
              "film_siren_fwd_gather: bad volume");
  CNG_REQUIRE(static_cast<long long>(D) * H * W < (1LL << 28), CNG_ERR_UNSUPPORTED, "film_siren_fwd_gather: volume exceeds the gather's 32-bit offsets");
  CNG_REQUIRE(vol_item_stride == 0 || vol_item_stride == static_cast<long long>(C) * D * H * W, CNG_ERR_INVALID_ARGUMENT,
              "film_siren_fwd_gather: vol_item_stride must be 0 (shared volume) or C*D*H*W");
  CNG_REQUIRE(precision == CNG_PREC_BF16 || precision == CNG_PREC_FP16, CNG_ERR_UNSUPPORTED, "film_siren_fwd_gather: tensor-core precisions only");
  for (int l = 0; l < L && l < 16; ++l)
    CNG_REQUIRE(layer_w_host[l] && layer_b_host[l], CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_gather: NULL layer %d", l);
  if (int e = cng_device_check()) return e;
  const cng::GatherSource src{vol_ndhwc, vol_item_stride, D, H, W, points};
  return cng::film_siren_tc_launch(nullptr, B, N, C, HID, L, layer_w_host, layer_b_host, freq, phase, final_w, final_b, sigmoid_rgb,
                                   precision == CNG_PREC_FP16 ? 1 : 0, workspace, workspace_bytes, rgb_sigma, nullptr, nullptr,
                                   cng::as_stream(stream), 0, 0, nullptr, nullptr, &src);
}

size_t cng_film_siren_res_scratch_bytes(void) {
  if (cng_device_check() != CNG_OK) return 0;
  return static_cast<size_t>(cng::sm_count()) * 2 * cng::kResScratchPerSlot;
}

int cng_film_siren_fwd_res(const float* feat, int B, long long N, int C, int HID, int L, const float* const* layer_w_host,
                           const float* const* layer_b_host, const float* freq, const float* phase, const float* final_w,
                           const float* final_b, int sigmoid_rgb, int precision, unsigned res_save_mask, unsigned res_add_mask,
                           void* workspace, size_t workspace_bytes, void* res_scratch, size_t res_scratch_bytes, float* rgb_sigma,
                           cng_stream_t stream) {
  CNG_REQUIRE(B >= 0 && N >= 0 && C >= 1 && HID >= 1 && L >= 1, CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_res: bad shape");
  if (B == 0 || N == 0) return CNG_OK;
  CNG_REQUIRE(feat && layer_w_host && layer_b_host && freq && phase && final_w && final_b && rgb_sigma,
              CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_res: NULL pointer");
  for (int l = 0; l < L && l < 16; ++l)
    CNG_REQUIRE(layer_w_host[l] && layer_b_host[l], CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_res: NULL layer %d", l);
  CNG_REQUIRE(L <= 16 && ((res_save_mask | res_add_mask) >> L) == 0, CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_res: mask bit beyond layer %d", L - 1);
  if (int e = cng_device_check()) return e;
  if (precision == CNG_PREC_FP32)
    return cng::film_siren_simt_launch(feat, B, N, C, HID, L, layer_w_host, layer_b_host, freq, phase, final_w, final_b,
                                       sigmoid_rgb, rgb_sigma, cng::as_stream(stream), res_save_mask, res_add_mask);
  CNG_REQUIRE(precision == CNG_PREC_BF16 || precision == CNG_PREC_FP16, CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_res: unknown precision %d", precision);
  CNG_REQUIRE((res_save_mask | res_add_mask) == 0 || (res_scratch != nullptr && res_scratch_bytes >= cng_film_siren_res_scratch_bytes() &&
                                                      (reinterpret_cast<uintptr_t>(res_scratch) & 15) == 0),
              CNG_ERR_WORKSPACE, "film_siren_fwd_res: residual scratch %zu < %zu bytes (or misaligned)", res_scratch_bytes,
              cng_film_siren_res_scratch_bytes());
  return cng::film_siren_tc_launch(feat, B, N, C, HID, L, layer_w_host, layer_b_host, freq, phase, final_w, final_b, sigmoid_rgb,
                                   precision == CNG_PREC_FP16 ? 1 : 0, workspace, workspace_bytes, rgb_sigma, nullptr, nullptr,
                                   cng::as_stream(stream), res_save_mask, res_add_mask, static_cast<float*>(res_scratch));
}

int cng_film_siren_fwd_train(const float* feat, int B, long long N, int C, int HID, int L, const float* const* layer_w_host,
                             const float* const* layer_b_host, const float* freq, const float* phase, const float* final_w,
                             const float* final_b, int sigmoid_rgb, int precision, unsigned res_save_mask, unsigned res_add_mask,
                             void* workspace, size_t workspace_bytes, void* res_scratch, size_t res_scratch_bytes, float* rgb_sigma,
                             void* x_dump, void* g_dump, void* feat_dump, cng_stream_t stream) {
  CNG_REQUIRE(B >= 0 && N >= 0 && C >= 1 && HID >= 1 && L >= 1, CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_train: bad shape");
  if (B == 0 || N == 0) return CNG_OK;
  CNG_REQUIRE(feat && layer_w_host && layer_b_host && freq && phase && final_w && final_b && rgb_sigma && x_dump && g_dump && feat_dump,
              CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_train: NULL pointer");
  CNG_REQUIRE(precision == CNG_PREC_BF16 || precision == CNG_PREC_FP16, CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_train: precision must be bf16 or fp16");
  for (int l = 0; l < L && l < 16; ++l)
    CNG_REQUIRE(layer_w_host[l] && layer_b_host[l], CNG_ERR_INVALID_ARGUMENT, "film_siren_fwd_train: NULL layer %d", l);
  if (int e = cng_device_check()) return e;
  CNG_REQUIRE((res_save_mask | res_add_mask) == 0 || (res_scratch != nullptr && res_scratch_bytes >= cng_film_siren_res_scratch_bytes()),
              CNG_ERR_WORKSPACE, "film_siren_fwd_train: residual scratch %zu < %zu bytes", res_scratch_bytes, cng_film_siren_res_scratch_bytes());
  return cng::film_siren_tc_launch(feat, B, N, C, HID, L, layer_w_host, layer_b_host, freq, phase, final_w, final_b, sigmoid_rgb,
                                   precision == CNG_PREC_FP16 ? 1 : 0, workspace, workspace_bytes, rgb_sigma, x_dump, g_dump, cng::as_stream(stream),
                                   res_save_mask, res_add_mask, static_cast<float*>(res_scratch), feat_dump);
}

}  // extern "C"
