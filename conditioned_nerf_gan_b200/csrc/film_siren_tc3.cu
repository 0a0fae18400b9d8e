// K2, layer-pipelined version: the FiLM-SIREN MLP on tcgen05 with ONE 128-point tile per CTA and a
// double-buffered TMEM accumulator, so that the MMAs of layer l+1 run while the epilogue of layer l is still
// producing their A operand.
//
// Why (measured on film_siren_tc.cu, profiles/r1f_tc_timeline.txt): with two tiles ping-ponging on one accumulator
// each, a tile's chain is MMA (2600 cycles) -> epilogue (3400-4100 when both tiles' epilogues share the MUFU unit)
// -> MMA ..., 6700-6800 cycles per pair of tile-layers although the tensor pipe and the MUFU unit each need only
// 2 x 2048.  The dependency that matters is finer than a layer: k-step j of layer l+1 needs only columns
// [16j, 16j+16) of layer l's output.  Here the 16 epilogue warps all work on the same tile, column sub-block by
// sub-block (4 * w columns: every warp takes w columns of its 32 rows, w from the schedule below), and signal each sub-block on its own
// mbarrier; the MMA warp issues the k-steps of the NEXT layer as the sub-blocks land, into the OTHER accumulator
// (TMEM: 2 x 256 fp32 columns).  The tensor pipe is therefore busy during the epilogue and only the last sub-block's
// k-steps plus the commit -> epilogue hand-off are exposed per layer.
//
//   group sequence per tile:  layer 0 .. layer L-1, head;   accumulator of group g (global count) = g & 1
//   E(l) (epilogue of layer l) reads accumulator g&1, adds the FiLM shift (rows of the current item in shared
//   memory), takes the sine and writes the 16-bit operand of layer l+1.
//   The next tile's features are fetched during E(L-1) into a separate 16 KB operand block, so its layer 0 runs
//   right behind the current tile's head.
//
// Shared memory: A tile 64 KB + layer-0 operand 16 KB + weight ring 4 x 32 KB + barriers + shift rows 8 KB = 217 KB.
// Same operand images, fold kernel and arithmetic order as film_siren_tc.cu: the two give bit-identical results
// (tests/test_gpu_parity.py).  Measured 10 % slower than that kernel (DESIGN.md 5 explains where the time goes), so it
// is the alternative (CNG_TC_V=3), not the default; film_siren_tc.cu also serves the training-mode dumps, L > 8 and
// the residual-block variants.
#include <stdlib.h>

#include <utility>

#include "film_siren_tc_common.cuh"

#ifndef CNG_TC3_EXP
#define CNG_TC3_EXP 0         // bottleneck experiments only (WRONG results), bit flags: 1 no weight copies after the first tile, 2 no operand
                              // stores, 4 no sines, 16 no proxy fence, 32 the issuing warp does not wait for sub-blocks
#endif
#ifndef CNG_TC3_SCHED
#define CNG_TC3_SCHED 8       // epilogue sub-block schedule, see make_sched()
#endif

namespace cng {
namespace tc3 {

constexpr int kRing = 4;
constexpr int kEpiWarps = 16;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kProducerWarp = kMmaWarp + 1;
constexpr int kNumThreads = 32 * (kProducerWarp + 1);
constexpr uint32_t kSmemA = 0;                                        // [4 K-blocks][128][64] 16-bit, 128B swizzle
constexpr uint32_t kSmemA0 = kATileBytes;                             // [128][64]: [x_hi(32) | x_lo(32)] of the next tile
constexpr uint32_t kSmemW = kSmemA0 + kABlockBytes;                   // 81920
constexpr uint32_t kSmemBar = kSmemW + kRing * kChunkBytes;           // 212992
constexpr uint32_t kSmemTab = kSmemBar + 512;                         // FiLM shift rows of the current item: [8][256] fp32
constexpr int kMaxL = 8;
constexpr uint32_t kSmemTotal = kSmemTab + kMaxL * kHID * 4;          // 221696

template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&v)[N]) {
  static_assert(N == 4 || N == 8 || N == 16, "columns per warp per sub-block");
  if constexpr (N == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
  } else if constexpr (N == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
  } else {
    CNG_TMEM_LD_16(taddr, v);
  }
}
template <int W>
__device__ __forceinline__ void tmem_ld_w(uint32_t taddr, uint32_t (&v)[16]) {      // W columns into the front of a 16-register buffer
  static_assert(W == 4 || W == 8 || W == 16, "columns per warp per step");
  if constexpr (W == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
  } else if constexpr (W == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
  } else {
    CNG_TMEM_LD_16(taddr, v);
  }
}
// Sub-block schedule of a layer's epilogue: step s covers 4*w(s) accumulator columns (every warp takes w(s) of them).
// Wide steps amortise the per-step chain (TMEM load wait -> sine -> pack -> store -> proxy fence -> arrive, ~300 cycles
// of latency per warp), narrow last steps keep short what is left for the tensor pipe after the epilogue's last arrive.
//   kSched 4 / 8 / 16: uniform width;  100: 16,16,16,8,8;  101: 16,16,16,8,4,4;  102: 16,16,8,8,8,8
struct Sched {
  int n;
  int w[16];
};
__host__ __device__ constexpr Sched make_sched(int k) {
  Sched r{};
  if (k == 100) { r.n = 5; const int w[5] = {16, 16, 16, 8, 8}; for (int i = 0; i < 5; ++i) r.w[i] = w[i]; }
  else if (k == 101) { r.n = 6; const int w[6] = {16, 16, 16, 8, 4, 4}; for (int i = 0; i < 6; ++i) r.w[i] = w[i]; }
  else if (k == 102) { r.n = 6; const int w[6] = {16, 16, 8, 8, 8, 8}; for (int i = 0; i < 6; ++i) r.w[i] = w[i]; }
  else { r.n = kHID / (4 * k); for (int i = 0; i < r.n; ++i) r.w[i] = k; }
  return r;
}
__host__ __device__ constexpr int sched_col0(const Sched& sc, int s) {
  int c = 0;
  for (int i = 0; i < s; ++i) c += 4 * sc.w[i];
  return c;
}
template <int V>
struct IC { static constexpr int value = V; };
template <class F, int... I>
__device__ __forceinline__ void for_each_ic(F&& f, std::integer_sequence<int, I...>) { (f(IC<I>{}), ...); }

// kSched: the schedule above.  kPolyOneIn / kHalf as in film_siren_tc.cu.
template <int kSched, int kPolyOneIn, bool kHalf>
__global__ void __launch_bounds__(kNumThreads, 1) film_siren_tc3_kernel(TcParams p) {
  constexpr Sched kS = make_sched(kSched);
  constexpr int kNSub = kS.n;                       // sub-blocks (steps) per layer
  static_assert(sched_col0(kS, kS.n) == kHID, "schedule must cover the 256 columns");
  using Steps = std::make_integer_sequence<int, kNSub>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.L;
  const uint32_t bar0 = s_base + kSmemBar;
  auto w_full = [&](int s) { return bar0 + 8u * s; };                 // [kRing]
  auto w_empty = [&](int s) { return bar0 + 32u + 8u * s; };          // [kRing]
  auto acc_full = [&](int a) { return bar0 + 64u + 8u * a; };         // [2]
  auto sub_ready = [&](int j) { return bar0 + 128u + 8u * j; };       // [kNSub <= 16]
  const uint32_t a0_ready = bar0 + 80u;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 96);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRing; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int a = 0; a < 2; ++a) mbar_init(acc_full(a), 1);
    mbar_init(a0_ready, kEpiWarps);
    for (int j = 0; j < kNSub; ++j) mbar_init(sub_ready(j), kEpiWarps);
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
    const bool elected = elect_one();
    int slot = 0;
    uint32_t phase = 0;
    for (long long t = first; t < p.total_tiles; t += G) {
      const uint8_t* img = p.images + static_cast<size_t>(t / p.tiles_per_item) * item_image_bytes(L);
      for (int l = 0; l <= L; ++l) {
        const int nchunks = (l == 0) ? 2 : (l < L ? 4 : 1);
        const uint32_t bytes = (l < L) ? kChunkBytes : kHeadBytes;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(w_empty(slot), phase ^ 1);
          if (elected) {
            if ((CNG_TC3_EXP & 1) && t != first) mbar_arrive(w_full(slot));
            else {
              mbar_arrive_expect_tx(w_full(slot), bytes);
              bulk_g2s(s_base + kSmemW + slot * kChunkBytes, img + chunk_offset(L, l, c), bytes, w_full(slot));
            }
          }
          __syncwarp();
          if (++slot == kRing) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer (warp-uniform loop, one elected lane issues) ===========================
    const bool elected = elect_one();
    int slot = 0;
    uint32_t phase = 0, sub_phase = 0, a0_phase = 0;
    uint32_t g = 0;                                  // global MMA-group counter: accumulator = g & 1
    constexpr uint32_t idesc_main = make_idesc(128, 256, kHalf);
    constexpr uint32_t idesc_head = make_idesc(128, 16, kHalf);
    const uint64_t a_tile_desc = make_desc(s_base + kSmemA);
    const uint64_t a0_desc = make_desc(s_base + kSmemA0);
    int iter = 0;
    for (long long t = first; t < p.total_tiles; t += G, ++iter) {
      // ---- layer 0: operand block A0 (staged by the epilogue warps during the previous tile's last layer) ----
      {
        const uint32_t d_tmem = tmem_base + (g & 1u) * kHID;
        mbar_wait(a0_ready, a0_phase);
        a0_phase ^= 1;
        tc_fence_after();
        if (elected) trace_event(p.trace, iter, 0, 0, 0);
        for (int c = 0; c < 2; ++c) {
          mbar_wait(w_full(slot), phase);
          tc_fence_after();
          const uint64_t b_desc = make_desc(s_base + kSmemW + slot * kChunkBytes);
          if (elected) {
            // chunk 0 = 4 k-steps over [x_hi|x_lo] x [W_hi|W_hi], chunk 1 = 2 k-steps over [x_hi] x [W_lo]
            tc_mma_bf16(d_tmem, a0_desc, b_desc, idesc_main, c ? 1u : 0u);
            tc_mma_bf16(d_tmem, a0_desc + 2, b_desc + 2, idesc_main, 1u);
            if (c == 0) {
              tc_mma_bf16(d_tmem, a0_desc + 4, b_desc + 4, idesc_main, 1u);
              tc_mma_bf16(d_tmem, a0_desc + 6, b_desc + 6, idesc_main, 1u);
            }
            tc_commit(w_empty(slot));
          }
          __syncwarp();
          if (++slot == kRing) { slot = 0; phase ^= 1; }
        }
        if (elected) {
          tc_commit(acc_full(g & 1u));
          trace_event(p.trace, iter, 0, 0, 1);
        }
        __syncwarp();
        ++g;
      }
      // ---- layers 1 .. L-1 and the head (l == L): k-steps follow the previous layer's epilogue sub-block by sub-block ----
      for (int l = 1; l <= L; ++l) {
        const uint32_t d_tmem = tmem_base + (g & 1u) * kHID;
        const bool head = (l == L);
        // every step lies inside one 64-wide K-block (= one weight ring slot for a hidden layer)
        if (!head) {
          for_each_ic([&](auto S) {
            constexpr int st = decltype(S)::value;
            constexpr int col0 = sched_col0(kS, st), k_first = col0 / 16, k_end = (col0 + 4 * kS.w[st]) / 16;
            constexpr bool opens = (k_first % 4 == 0), closes = (k_end % 4 == 0);
            if (opens) mbar_wait_lean(w_full(slot), phase);          // weights land a layer ahead
            if (!(CNG_TC3_EXP & 32) || st == 0) mbar_wait_lean(sub_ready(st), sub_phase);
            if (st == 0 && elected) trace_event(p.trace, iter, l, 0, 0);
            tc_fence_after();
            if (elected) {
              const uint64_t b_desc = make_desc(s_base + kSmemW + slot * kChunkBytes);
#pragma unroll
              for (int k16 = k_first; k16 < k_end; ++k16)
                tc_mma_bf16(d_tmem, a_tile_desc + (k16 >> 2) * (kABlockBytes >> 4) + 2 * (k16 & 3), b_desc + 2 * (k16 & 3), idesc_main, k16 ? 1u : 0u);
              if (closes) tc_commit(w_empty(slot));
            }
            if (closes) { if (++slot == kRing) { slot = 0; phase ^= 1; } }
          }, Steps{});
        } else {
          const uint64_t b_desc = make_desc(s_base + kSmemW + slot * kChunkBytes);
          mbar_wait_lean(w_full(slot), phase);
          for_each_ic([&](auto S) {
            constexpr int st = decltype(S)::value;
            constexpr int col0 = sched_col0(kS, st), k_first = col0 / 16, k_end = (col0 + 4 * kS.w[st]) / 16;
            mbar_wait_lean(sub_ready(st), sub_phase);
            tc_fence_after();
            if (elected) {
#pragma unroll
              for (int k16 = k_first; k16 < k_end; ++k16)
                tc_mma_bf16(d_tmem, a_tile_desc + (k16 >> 2) * (kABlockBytes >> 4) + 2 * (k16 & 3), b_desc + (k16 >> 2) * (2048 >> 4) + 2 * (k16 & 3),
                            idesc_head, k16 ? 1u : 0u);
              if (st == kNSub - 1) tc_commit(w_empty(slot));
            }
          }, Steps{});
          if (++slot == kRing) { slot = 0; phase ^= 1; }
        }
        sub_phase ^= 1;
        if (elected) {
          tc_commit(acc_full(g & 1u));
          trace_event(p.trace, iter, l, 0, 1);
        }
        ++g;
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    const int q = warp & 3;                       // TMEM lane quarter == warp_id % 4
    const int cg = warp >> 2;                     // column group within a sub-block
    const int row = q * 32 + lane;
    const int tid = threadIdx.x;                  // 0..511
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t acc_phase = 0;                       // bit a = parity of acc_full(a)
    uint32_t g = 0;
    const bool tracer = warp == 0 && lane == 0;

    // features of tile `t` -> registers (issue early), then -> A0 block as [x_hi(32) | x_lo(32)]
    float4 fv[2];
    auto feat_load = [&](long long t) {
      const TileInfo ti = tile_info(p, t);
      const float4* f = reinterpret_cast<const float4*>(p.feat + (static_cast<size_t>(ti.item) * p.N + ti.n0) * kC0);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = tid + 512 * i;
        fv[i] = (idx >> 3) < ti.rows ? __ldg(f + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto feat_store = [&]() {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = tid + 512 * i;
        const int r = idx >> 3, c4 = idx & 7;
        const float4 v = fv[i];
        uint2 hi, lo;
        hi.x = pack2<kHalf>(v.x, v.y); hi.y = pack2<kHalf>(v.z, v.w);
        lo.x = pack2<kHalf>(v.x - from16<kHalf>(to16<kHalf>(v.x)), v.y - from16<kHalf>(to16<kHalf>(v.y)));
        lo.y = pack2<kHalf>(v.z - from16<kHalf>(to16<kHalf>(v.z)), v.w - from16<kHalf>(to16<kHalf>(v.w)));
        *reinterpret_cast<uint2*>(smem + kSmemA0 + sw128_offset(r, 4 * c4)) = hi;
        *reinterpret_cast<uint2*>(smem + kSmemA0 + sw128_offset(r, 32 + 4 * c4)) = lo;
      }
    };
    auto signal = [&](uint32_t bar) {
      if (!(CNG_TC3_EXP & 16)) fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };

    // FiLM shift rows of the current batch item, [L][256] fp32 in shared memory (reloaded when the item changes:
    // every ~20 tiles at the bench shape); the epilogue adds them to the accumulator (one LDS.128 per 4 columns,
    // warp-broadcast, issued a sub-block ahead).  Measured alternatives: pre-storing the shift in the accumulator
    // (tcgen05.st, as film_siren_tc.cu does) cost +900 cycles per layer inside the sub-block loop and +850 as a
    // burst after it (L2 latency of the shift loads each time).
    float* tab = reinterpret_cast<float*>(smem + kSmemTab);
    int tab_item = -1;
    auto tab_load = [&](int item) {
      if (tid < L * (kHID / 4))
        reinterpret_cast<float4*>(tab)[tid] = __ldg(reinterpret_cast<const float4*>(p.shift + static_cast<size_t>(item) * L * kHID) + tid);
      tab_item = item;
      named_bar_sync(1, 32 * kEpiWarps);
    };

    // ---- prologue: first tile's features ----
    if (first < p.total_tiles) {
      feat_load(first);
      feat_store();
      signal(a0_ready);
    }

    int iter = 0;
    for (long long t = first; t < p.total_tiles; t += G, ++iter) {
      const TileInfo ti = tile_info(p, t);
      const bool has_next = t + G < p.total_tiles;
      // every warp is past its last read of the old rows here (it has waited for this tile's predecessor's head, which
      // needs every warp's last sub-block), so one barrier after the reload is enough
      if (ti.item != tab_item) tab_load(ti.item);
      for (int l = 0; l < L; ++l) {
        const bool stage_next = (l == L - 1) && has_next;
        if (stage_next) feat_load(t + G);                      // in flight during this layer's sines
        const uint32_t acc = g & 1u;
        const uint32_t t_acc = t_lane + acc * kHID;
        const float* sh_row = tab + l * kHID;
        float sh[2][16];
        uint32_t v[2][16];
        auto ld_step = [&](auto S) {                     // TMEM load + shift rows of step S into buffer S & 1 (asynchronous)
          constexpr int st = decltype(S)::value, W = kS.w[st], col0 = sched_col0(kS, st);
          tmem_ld_w<W>(t_acc + col0 + cg * W, v[st & 1]);
#pragma unroll
          for (int e = 0; e < W; e += 4)
            *reinterpret_cast<float4*>(&sh[st & 1][e]) = *reinterpret_cast<const float4*>(sh_row + col0 + cg * W + e);
        };
        mbar_wait(acc_full(acc), (acc_phase >> acc) & 1u);
        acc_phase ^= 1u << acc;
        tc_fence_after();
        if (tracer) trace_event(p.trace, iter, l, 0, 2);
        ld_step(IC<0>{});
        for_each_ic([&](auto S) {
          constexpr int st = decltype(S)::value, W = kS.w[st], b = st & 1;
          tmem_ld_wait();
          if constexpr (st + 1 < kNSub) ld_step(IC<st + 1>{});
          const int col = sched_col0(kS, st) + cg * W;
          const int kb = col >> 6, k = col & 63;
          uint8_t* dst = smem + kSmemA + kb * kABlockBytes + row * 128;
          float sn[W];
#pragma unroll
          for (int e = 0; e < W; ++e) {
            const float x = __uint_as_float(v[b][e]) + sh[b][e];
            if (kPolyOneIn > 0 && (e % (kPolyOneIn > 0 ? kPolyOneIn : 1)) == kPolyOneIn - 1) sn[e] = sin_fma(x);
            else if (CNG_TC3_EXP & 4) sn[e] = x * 0.001f;
            else sn[e] = __sinf(x);
          }
          uint32_t o[W / 2];
#pragma unroll
          for (int e = 0; e < W; e += 2) o[e / 2] = pack2<kHalf>(sn[e], sn[e + 1]);
          if ((CNG_TC3_EXP & 2) && o[0] != 0x12345678u) {
          } else if constexpr (W == 4) {
            *reinterpret_cast<uint2*>(dst + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2) = make_uint2(o[0], o[1]);
          } else {
#pragma unroll
            for (int c = 0; c < W / 8; ++c)
              *reinterpret_cast<uint4*>(dst + ((((k >> 3) + c) ^ (row & 7)) << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
          }
          if (st == kNSub - 1) {
            if (stage_next) feat_store();
            tc_fence_before();          // this warp's accumulator reads are ordered before the MMAs that overwrite it (two groups on)
            if (tracer) trace_event(p.trace, iter, l, 0, 3);
          }
          signal(sub_ready(st));
        }, Steps{});
        if (stage_next) signal(a0_ready);
        ++g;
      }
      // ---- head: 4 accumulator columns -> bias, sigmoid(rgb), store (column group 0 holds them) ----
      {
        const uint32_t acc = g & 1u;
        mbar_wait(acc_full(acc), (acc_phase >> acc) & 1u);
        acc_phase ^= 1u << acc;
        tc_fence_after();
        if (cg == 0) {
          uint32_t r[4];
          tmem_ld_n<4>(t_lane + acc * kHID, r);
          tmem_ld_wait();
          float4 o;
          const float4 fb = __ldg(reinterpret_cast<const float4*>(p.final_b));
          o.x = __uint_as_float(r[0]) + fb.x;
          o.y = __uint_as_float(r[1]) + fb.y;
          o.z = __uint_as_float(r[2]) + fb.z;
          o.w = __uint_as_float(r[3]) + fb.w;
          if (p.sigmoid_rgb) {
            o.x = 1.f / (1.f + __expf(-o.x));
            o.y = 1.f / (1.f + __expf(-o.y));
            o.z = 1.f / (1.f + __expf(-o.z));
          }
          if (row < ti.rows) reinterpret_cast<float4*>(p.out)[static_cast<size_t>(ti.item) * p.N + ti.n0 + row] = o;
        }
        tc_fence_before();
        ++g;
      }
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

}  // namespace tc3

// p.images / p.shift: the fold kernel's output (film_siren_tc.cu); inference only (no activation dumps)
int film_siren_tc3_launch(TcParams p, int poly, cudaStream_t stream) {
  using KernelFn = void (*)(TcParams);
  constexpr int CW = CNG_TC3_SCHED;
  KernelFn fn;
  if (p.half_operands) fn = poly == 4 ? tc3::film_siren_tc3_kernel<CW, 4, true> : poly == 8 ? tc3::film_siren_tc3_kernel<CW, 8, true> : tc3::film_siren_tc3_kernel<CW, 0, true>;
  else fn = poly == 4 ? tc3::film_siren_tc3_kernel<CW, 4, false> : poly == 8 ? tc3::film_siren_tc3_kernel<CW, 8, false> : tc3::film_siren_tc3_kernel<CW, 0, false>;
  cudaError_t ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(tc3::kSmemTotal));
  if (ce != cudaSuccess) return fail(static_cast<int>(ce), "film_siren_fwd(tc3): smem attribute: %s", cudaGetErrorString(ce));
  const long long grid = min(static_cast<long long>(sm_count()), p.total_tiles);
  fn<<<static_cast<unsigned>(grid), tc3::kNumThreads, tc3::kSmemTotal, stream>>>(p);
  return check_launch("cng_film_siren_fwd(tc3)");
}

}  // namespace cng
