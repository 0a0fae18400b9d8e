// K2, CTA-pair version: the FiLM-SIREN MLP on tcgen05 with cta_group::2 (UMMA 256 x 256 x 16 across two SMs).
//
// Why pairs: with one CTA per SM every 128x256x16 MMA reads 4 KB of A and 8 KB of B from shared memory while
// the weight ring is refilled at 8 KB per MMA and the epilogue writes 4 KB of activations per MMA -- 24 KB per
// 128 tensor cycles against a 128 B/clk shared-memory port, so the tensor pipe idles ~55 % of the time
// (profiles/r1a_film_siren_tc_raw.txt: tensor 44 %, MUFU 49 %, both far from their limits).  In a pair each
// CTA stages only HALF of every weight block (N rows [128r, 128r+128)) and its tensor core reads the other
// half from the peer: B traffic and ring refill halve (16 KB per MMA), and L2 -> SMEM weight traffic halves.
//
// Layout per CTA (rank r = %cluster_ctarank): rows [128r, 128r+128) of two 256-point super-tiles ("slots")
// live in this CTA's shared memory (A operand) and TMEM (accumulator, 2 x 256 columns).  Roles per CTA:
//   warps 0-7 / 8-15   epilogue of slot 0 / 1 (as in film_siren_tc.cu: sin on the fp32 accumulator, bf16 A
//                      operand of the next layer into shared memory, next layer's FiLM shift into TMEM)
//   warp 16            rank 0: MMA issuer (one lane issues tcgen05.mma.cta_group::2 for both CTAs)
//                      rank 1: relay -- forwards "my half of weight block k has landed" to rank 0
//   warp 17            weight producer: cp.async.bulk of this CTA's half blocks into a 6 x 16 KB ring
// Barriers: act_ready[slot] lives in rank 0 and counts one elected arrive per epilogue warp of BOTH CTAs
// (remote mbarrier.arrive.release.cluster); w_empty[] / acc_full[] are signalled in both CTAs at once by
// tcgen05.commit ... .multicast::cluster.
#include <stdlib.h>

#include "film_siren_tc_common.cuh"

namespace cng {
namespace tc2 {

constexpr int kRing = 6;
constexpr int kHalfChunk = kChunkBytes / 2;          // [128 n][64 k] bf16 = 16 KB
constexpr int kEpiWarpsPerSlot = 8;
constexpr int kMmaWarp = 2 * kEpiWarpsPerSlot;
constexpr int kProducerWarp = kMmaWarp + 1;
constexpr int kNumThreads = 32 * (kProducerWarp + 1);
constexpr int kSuperM = 2 * kTileM;                  // points per super-tile
constexpr uint32_t kSmemA = 0;
constexpr uint32_t kSmemW = 2 * kATileBytes;                         // 131072
constexpr uint32_t kSmemBar = kSmemW + kRing * kHalfChunk;           // 229376
constexpr uint32_t kSmemTotal = kSmemBar + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait that also acquires at cluster scope (the arrivals come from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) break;
    if ((it & 63u) == 63u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {      // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void tc_mma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// p.tiles_per_item / p.total_tiles count 256-point super-tiles here
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1) film_siren_tc2_kernel(TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t s_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int L = p.L;
  const uint32_t bar0 = s_base + kSmemBar;
  auto w_full = [&](int s) { return bar0 + 8u * s; };                 // local: this CTA's half block landed
  auto w_empty = [&](int s) { return bar0 + 48u + 8u * s; };          // local: ring slot free again (pair commit)
  auto w_peer = [&](int s) { return bar0 + 96u + 8u * s; };           // rank 0: the peer's half block landed
  auto act_ready = [&](int x) { return bar0 + 144u + 8u * x; };       // rank 0: operands of slot x ready in both CTAs
  auto acc_full = [&](int x) { return bar0 + 160u + 8u * x; };        // local: accumulator of slot x complete (pair commit)
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 192);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kRing; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); mbar_init(w_peer(s), 1); }
    for (int x = 0; x < 2; ++x) { mbar_init(act_ready(x), 2 * kEpiWarpsPerSlot); mbar_init(acc_full(x), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_base + kSmemBar + 192), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const long long G = gridDim.x >> 1;                 // CTA pairs
  const long long first = blockIdx.x >> 1;

  if (warp == kProducerWarp) {
    // =========================== weight producer (this CTA's half of every block) ===========================
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
          for (int x = 0; x < nx; ++x) {
            const uint8_t* img = p.images + static_cast<size_t>(x == 0 ? item0 : item1) * item_image_bytes(L);
            for (int c = 0; c < nchunks; ++c) {
              mbar_wait(w_empty(slot), phase ^ 1);
              const uint32_t dst = s_base + kSmemW + slot * kHalfChunk;
              if (elected) {
                if (l < L) {
                  mbar_arrive_expect_tx(w_full(slot), kHalfChunk);
                  bulk_g2s(dst, img + chunk_offset(L, l, c) + rank * kHalfChunk, kHalfChunk, w_full(slot));
                } else {
                  // head: 4 K-blocks of [16 n][64 k]; this CTA takes rows [8r, 8r+8) of each (1 KB)
                  mbar_arrive_expect_tx(w_full(slot), 4096);
#pragma unroll
                  for (int kb = 0; kb < 4; ++kb)
                    bulk_g2s(dst + kb * 1024, img + chunk_offset(L, L, 0) + kb * 2048 + rank * 1024, 1024, w_full(slot));
                }
              }
              __syncwarp();
              if (++slot == kRing) { slot = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    const bool elected = elect_one();
    int slot = 0;
    uint32_t phase = 0;
    if (rank == 0) {
      // =========================== MMA issuer (for both CTAs); warp-uniform loop, one elected lane issues ===========================
      uint32_t act_phase = 0;
      constexpr uint32_t idesc_main = make_idesc(256, 256);
      constexpr uint32_t idesc_head = make_idesc(256, 16);
      int iter = 0;
      for (long long t0 = first; t0 < p.total_tiles; t0 += 2 * G, ++iter) {
        const int nx = (t0 + G < p.total_tiles) ? 2 : 1;
        for (int l = 0; l <= L; ++l) {
          const int nchunks = (l == 0) ? 2 : (l < L ? 4 : 1);
          for (int x = 0; x < nx; ++x) {
            mbar_wait_cluster(act_ready(x), (act_phase >> x) & 1u);
            act_phase ^= 1u << x;
            tc_fence_after();
            if (elected) trace_event(p.trace, iter, l, x, 0);
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(x) * kHID;
            const uint64_t a_desc0 = make_desc(s_base + kSmemA + x * kATileBytes);
            for (int c = 0; c < nchunks; ++c) {
              mbar_wait(w_full(slot), phase);
              mbar_wait_cluster(w_peer(slot), phase);
              tc_fence_after();
              const uint64_t b_desc = make_desc(s_base + kSmemW + slot * kHalfChunk);
              if (elected) {
                if (l < L) {
                  const uint64_t a_desc = a_desc0 + (l == 0 ? 0 : c * (kABlockBytes >> 4));
                  tc_mma2_bf16(d_tmem, a_desc, b_desc, idesc_main, 1u);               // D holds the shift: always accumulate
                  tc_mma2_bf16(d_tmem, a_desc + 2, b_desc + 2, idesc_main, 1u);
                  if (!(l == 0 && c == 1)) {
                    tc_mma2_bf16(d_tmem, a_desc + 4, b_desc + 4, idesc_main, 1u);
                    tc_mma2_bf16(d_tmem, a_desc + 6, b_desc + 6, idesc_main, 1u);
                  }
                } else {
#pragma unroll
                  for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                      tc_mma2_bf16(d_tmem, a_desc0 + kb * (kABlockBytes >> 4) + 2 * ks, b_desc + kb * (1024 >> 4) + 2 * ks, idesc_head,
                                   (kb | ks) ? 1u : 0u);
                }
                tc_commit_pair(w_empty(slot));
              }
              __syncwarp();
              if (++slot == kRing) { slot = 0; phase ^= 1; }
            }
            if (elected) {
              tc_commit_pair(acc_full(x));
              trace_event(p.trace, iter, l, x, 1);
            }
            __syncwarp();
          }
        }
      }
    } else {
      // =========================== relay: my half block k is in shared memory -> tell rank 0 ===========================
      const uint32_t peer_bar0 = map_to_cta(w_peer(0), 0);
      for (long long t0 = first; t0 < p.total_tiles; t0 += 2 * G) {
        const int nx = (t0 + G < p.total_tiles) ? 2 : 1;
        for (int l = 0; l <= L; ++l) {
          const int nchunks = ((l == 0) ? 2 : (l < L ? 4 : 1)) * nx;
          for (int c = 0; c < nchunks; ++c) {
            mbar_wait(w_full(slot), phase);
            if (elected) mbar_arrive_cluster(peer_bar0 + 8u * slot);
            __syncwarp();
            if (++slot == kRing) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // =========================== epilogue warps (slot x = warp / 8) ===========================
    const int x = warp / kEpiWarpsPerSlot;
    const int q = warp & 3;                       // TMEM lane quarter == warp_id % 4
    const int half = (warp >> 2) & 1;             // accumulator columns [128*half, 128*half + 128)
    const int row = q * 32 + lane;
    const uint32_t a_base = kSmemA + x * kATileBytes;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(x) * kHID;
    const uint32_t ready_remote = map_to_cta(act_ready(x), 0);
    uint32_t acc_phase = 0;
    int iter = 0;
    const bool tracer = (warp % kEpiWarpsPerSlot) == 0 && lane == 0;
    auto signal_ready = [&]() {
      tmem_st_wait();
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(ready_remote);
    };
    for (long long t = first + x * G; t < p.total_tiles; t += 2 * G, ++iter) {
      const int item = static_cast<int>(t / p.tiles_per_item);
      const long long n0 = (t - static_cast<long long>(item) * p.tiles_per_item) * kSuperM + rank * kTileM;
      const int rows = static_cast<int>(max(0LL, min(static_cast<long long>(kTileM), p.N - n0)));
      const float* shift_item = p.shift + static_cast<size_t>(item) * L * kHID;
      Shift32 sh;
#pragma unroll 1
      for (int cc = 4 * half; cc < 4 * half + 4; ++cc) {
        sh.load(shift_item + cc * 32);
        sh.store(t_lane + cc * 32);
      }
      {
        const float4* f = reinterpret_cast<const float4*>(p.feat + (static_cast<size_t>(item) * p.N + n0) * kC0);
#pragma unroll
        for (int it = 4 * half; it < 4 * half + 4; ++it) {
          const int r = q * 32 + it * 4 + (lane >> 3);
          const int c4 = lane & 7;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < rows) v = __ldg(f + r * 8 + c4);
          const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z),
                              h3 = __float2bfloat16_rn(v.w);
          uint2 hi, lo;
          hi.x = pack_bf16(v.x, v.y); hi.y = pack_bf16(v.z, v.w);
          lo.x = pack_bf16(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
          lo.y = pack_bf16(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
          *reinterpret_cast<uint2*>(smem + a_base + sw128_offset(r, 4 * c4)) = hi;
          *reinterpret_cast<uint2*>(smem + a_base + sw128_offset(r, 32 + 4 * c4)) = lo;
        }
      }
      signal_ready();
      for (int l = 0; l < L; ++l) {
        const bool more = l + 1 < L;
        const float* shift_next = shift_item + (more ? l + 1 : l) * kHID;
        sh.load(shift_next + 4 * half * 32);
        mbar_wait(acc_full(x), acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (tracer) trace_event(p.trace, iter, l, x, 2);
#pragma unroll 1
        for (int cc = 4 * half; cc < 4 * half + 4; ++cc) {
          uint32_t v[32];
          CNG_TMEM_LD_32(t_lane + cc * 32, v);
          tmem_ld_wait();
          if (more) sh.store(t_lane + cc * 32);
          if (cc + 1 < 4 * half + 4) sh.load(shift_next + (cc + 1) * 32);
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) o[j / 2] = pack_bf16(__sinf(__uint_as_float(v[j])), __sinf(__uint_as_float(v[j + 1])));
          uint8_t* blk = smem + a_base + (cc >> 1) * kABlockBytes + row * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int chunk = ((cc & 1) * 4 + i) ^ (row & 7);
            *reinterpret_cast<uint4*>(blk + chunk * 16) = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
          }
        }
        if (tracer) trace_event(p.trace, iter, l, x, 3);
        signal_ready();
      }
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
        if (row < rows) reinterpret_cast<float4*>(p.out)[static_cast<size_t>(item) * p.N + n0 + row] = o;
      }
      tc_fence_before();
    }
  }
  // ---- teardown: both CTAs must be done with each other's shared memory / TMEM ----
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace tc2

// p.images / p.shift are the fold kernel's output (film_siren_tc.cu); tiles are recounted as 256-point super-tiles
int film_siren_tc2_launch(TcParams p, cudaStream_t stream) {
  p.tiles_per_item = (p.N + tc2::kSuperM - 1) / tc2::kSuperM;
  p.total_tiles = p.tiles_per_item * p.B;
  static bool attr_set[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fail(CNG_ERR_NO_DEVICE, "film_siren_fwd(bf16, pairs): no current device");
  if (!attr_set[dev]) {
    cudaError_t ce = cudaFuncSetAttribute(tc2::film_siren_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(tc2::kSmemTotal));
    if (ce != cudaSuccess) return fail(static_cast<int>(ce), "film_siren_fwd(bf16, pairs): smem attribute: %s", cudaGetErrorString(ce));
    attr_set[dev] = true;
  }
  const long long pairs = min(static_cast<long long>(sm_count() / 2), p.total_tiles);
  tc2::film_siren_tc2_kernel<<<static_cast<unsigned>(2 * pairs), tc2::kNumThreads, tc2::kSmemTotal, stream>>>(p);
  return check_launch("cng_film_siren_fwd(bf16, pairs)");
}

}  // namespace cng
