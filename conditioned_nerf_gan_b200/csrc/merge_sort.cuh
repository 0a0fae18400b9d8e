// Per-warp depth sort of the fine-first concatenation of a ray's samples (generators/generators.py:163-165):
// a bitonic network in shared memory on 64-bit keys (order-preserving bits of t in the high word, position in
// the concatenation in the low word), i.e. the stable order of the oracle (ties keep fine before coarse).
#pragma once
#include "cng_common.cuh"

namespace cng {

__device__ __forceinline__ uint32_t sortable_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_sortable_bits(uint32_t s) {
  return __uint_as_float((s & 0x80000000u) ? (s & 0x7fffffffu) : ~s);
}

// keys[0..n2): the first n entries valid, the rest padded by the caller with 0xffff....; n2 a power of two >= 32
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long* keys, int n2, int lane) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = lane; i < (n2 >> 1); i += 32) {
        // i-th compare-exchange of this stage: partner indices lo < hi differing in bit j
        const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
        const int hi = lo | j;
        const bool up = (lo & k) == 0;
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
      __syncwarp();
    }
  }
}

// The same network with the keys in registers: key index i = slot * 32 + lane.  Exchanges at distance >= 32 are
// register swaps inside a lane, smaller distances are warp shuffles; no shared-memory traffic and no __syncwarp.
template <int SLOTS>
__device__ __forceinline__ void warp_bitonic_sort_regs(unsigned long long (&key)[SLOTS], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * SLOTS; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int js = j >> 5;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          if ((s & js) == 0) {
            const bool up = ((s * 32) & k) == 0;                 // k >= 64 here: the direction depends on the slot only
            const unsigned long long a = key[s], b = key[s | js];
            if ((a > b) == up) { key[s] = b; key[s | js] = a; }
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          const unsigned long long a = key[s];
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, a, j);
          const bool up = (((s * 32 + lane) & k) == 0);
          const bool lower = (lane & j) == 0;
          const unsigned long long mn = a < o ? a : o, mx = a < o ? o : a;
          key[s] = (lower == up) ? mn : mx;
        }
      }
    }
  }
}

__host__ __device__ inline int next_pow2_min32(int n) {
  int p = 32;
  while (p < n) p <<= 1;
  return p;
}

// Register-sort variant of load_and_sort_ray for n2 == 32 * SLOTS; the sorted keys are written to keys[] once.
template <int SLOTS>
__device__ __forceinline__ void load_and_sort_ray_regs(unsigned long long* keys, const float* __restrict__ t_fine,
                                                       const float* __restrict__ t_coarse, long long ray, int S, int n, int lane) {
  const bool two = t_fine != nullptr;
  unsigned long long key[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int e = s * 32 + lane;
    key[s] = ~0ull;
    if (e < n) {
      const float te = two ? (e < S ? __ldg(t_fine + ray * S + e) : __ldg(t_coarse + ray * S + (e - S))) : __ldg(t_coarse + ray * S + e);
      key[s] = (static_cast<unsigned long long>(sortable_bits(te)) << 32) | static_cast<unsigned>(e);
    }
  }
  if (two) warp_bitonic_sort_regs<SLOTS>(key, lane);
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) keys[s * 32 + lane] = key[s];
  __syncwarp();
}

// Fast path of load_and_sort_ray when the coarse distances are already non-decreasing (they are: stratified jitter never
// crosses a neighbour): sort only the S fine keys in registers, then MERGE -- the rank of a key in the merged order is its
// rank in its own list plus the number of keys of the other list that precede it (binary search in shared memory).
// Keys are unique (t bits, source index), fine indices < coarse indices, so this is exactly the stable fine-first order.
// scratch: 2 * 32*SLOTS + ... 64-bit words after keys[0..n): fine list at keys + n2, coarse list at keys + n2 + 32*SLOTS.
template <int SLOTS>
__device__ __forceinline__ void sort_fine_and_merge(unsigned long long* keys, unsigned long long* fine_s, unsigned long long* coarse_s,
                                                    const float* __restrict__ t_fine, const float* __restrict__ t_coarse, long long ray,
                                                    int S, int lane) {
  unsigned long long key[SLOTS];
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int e = s * 32 + lane;
    key[s] = ~0ull;
    if (e < S) key[s] = (static_cast<unsigned long long>(sortable_bits(__ldg(t_fine + ray * S + e))) << 32) | static_cast<unsigned>(e);
  }
  warp_bitonic_sort_regs<SLOTS>(key, lane);
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) fine_s[s * 32 + lane] = key[s];
  for (int e = lane; e < S; e += 32)
    coarse_s[e] = (static_cast<unsigned long long>(sortable_bits(__ldg(t_coarse + ray * S + e))) << 32) | static_cast<unsigned>(S + e);
  __syncwarp();
  auto lower_bound = [&](const unsigned long long* a, unsigned long long k) {     // number of a[0..S) that are < k
    int lo = 0, hi = S;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (a[mid] < k) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  for (int e = lane; e < S; e += 32) {
    const unsigned long long kf = fine_s[e], kc = coarse_s[e];
    keys[e + lower_bound(coarse_s, kf)] = kf;
    keys[e + lower_bound(fine_s, kc)] = kc;
  }
  __syncwarp();
}

// true iff t_coarse[ray, 0..S) is non-decreasing (warp vote)
__device__ __forceinline__ bool coarse_is_sorted(const float* __restrict__ t_coarse, long long ray, int S, int lane) {
  bool ok = true;
  for (int e = lane; e + 1 < S; e += 32) ok = ok && (__ldg(t_coarse + ray * S + e) <= __ldg(t_coarse + ray * S + e + 1));
  return __all_sync(0xffffffffu, ok);
}

// 64-bit words of shared memory one warp needs for n samples (S per list)
__host__ __device__ inline int merge_smem_words(int n, int S) { return next_pow2_min32(n) + next_pow2_min32(S) + S; }

// Loads the ray's distances (fine first, then coarse; or coarse only), sorts, and leaves keys[s] = (t, source index).
__device__ __forceinline__ void load_and_sort_ray(unsigned long long* keys, const float* __restrict__ t_fine,
                                                  const float* __restrict__ t_coarse, long long ray, int S, int n, int n2, int lane) {
  const bool two = t_fine != nullptr;
  if (two && S <= 256 && coarse_is_sorted(t_coarse, ray, S, lane)) {
    unsigned long long* fine_s = keys + n2;
    unsigned long long* coarse_s = fine_s + next_pow2_min32(S);
    switch (next_pow2_min32(S)) {
      case 32: sort_fine_and_merge<1>(keys, fine_s, coarse_s, t_fine, t_coarse, ray, S, lane); return;
      case 64: sort_fine_and_merge<2>(keys, fine_s, coarse_s, t_fine, t_coarse, ray, S, lane); return;
      case 128: sort_fine_and_merge<4>(keys, fine_s, coarse_s, t_fine, t_coarse, ray, S, lane); return;
      default: sort_fine_and_merge<8>(keys, fine_s, coarse_s, t_fine, t_coarse, ray, S, lane); return;
    }
  }
  // up to 256 keys: sort in registers (1 to 8 keys per lane)
  switch (n2) {
    case 32: load_and_sort_ray_regs<1>(keys, t_fine, t_coarse, ray, S, n, lane); return;
    case 64: load_and_sort_ray_regs<2>(keys, t_fine, t_coarse, ray, S, n, lane); return;
    case 128: load_and_sort_ray_regs<4>(keys, t_fine, t_coarse, ray, S, n, lane); return;
    case 256: load_and_sort_ray_regs<8>(keys, t_fine, t_coarse, ray, S, n, lane); return;
    default: break;
  }
  for (int e = lane; e < n2; e += 32) {
    unsigned long long key = ~0ull;
    if (e < n) {
      const float te = two ? (e < S ? __ldg(t_fine + ray * S + e) : __ldg(t_coarse + ray * S + (e - S))) : __ldg(t_coarse + ray * S + e);
      key = (static_cast<unsigned long long>(sortable_bits(te)) << 32) | static_cast<unsigned>(e);
    }
    keys[e] = key;
  }
  __syncwarp();
  if (two) warp_bitonic_sort(keys, n2, lane);
}

__device__ __forceinline__ float key_t(unsigned long long k) { return from_sortable_bits(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ int key_src(unsigned long long k) { return static_cast<int>(k & 0xffffffffu); }

}  // namespace cng
