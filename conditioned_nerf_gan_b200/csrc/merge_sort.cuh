// Per-warp depth order of the fine-first concatenation of a ray's samples (generators/generators.py:163-165): the stable sort of
// [fine | coarse] by distance, ties keeping fine before coarse and lower source index first -- the order of the oracle.
//
// Result format: keys[s] for s in [0, n) = (IEEE bits of t) << 32 | source index in the concatenation (fine e, coarse S + e).
//
// Fast path (coarse distances non-decreasing, as stratified jitter always leaves them; <= 256 samples per list):
//   1. the S fine keys are sorted in REGISTERS by a bitonic network on lane-major keys (index = lane * SLOTS + slot: exchanges at
//      distance < SLOTS stay inside a lane, 15 stages shuffle).  When the ray's fine distances span fewer than 2^(32 - index bits)
//      representable floats -- they do unless the ray crosses many binades -- the network runs on 32-bit keys
//      ((sortable bits - min) << index bits | index), one SHFL and one compare per exchange instead of two and two;
//   2. a fine key's place in the merged order = its rank among the fine keys + the number of coarse distances strictly below it
//      (branch-free binary search over the coarse distances in shared memory);
//   3. the coarse keys fill the places the fine keys left empty, in order (an occupancy byte per place, a zero count per lane and
//      one warp scan): no second search.
// Anything else (unsorted coarse list, NaN) takes the generic path: a bitonic network over all n keys in shared memory.
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

__device__ __forceinline__ unsigned long long make_key(float t, int src) {
  return (static_cast<unsigned long long>(__float_as_uint(t)) << 32) | static_cast<unsigned>(src);
}
__device__ __forceinline__ float key_t(unsigned long long k) { return __uint_as_float(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ int key_src(unsigned long long k) { return static_cast<int>(k & 0xffffffffu); }

// keys[0..n2) of SORTABLE keys (sortable bits << 32 | index): the first n entries valid, the rest 0xffff....; n2 a power of two >= 32
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

// The network on LANE-MAJOR keys in registers: key index i = lane * SLOTS + slot.  Equal keys may only be padding.
template <int SLOTS, typename K>
__device__ __forceinline__ void warp_bitonic_sort_lane_major(K (&key)[SLOTS], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * SLOTS; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j < SLOTS) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          if ((s & j) == 0) {
            const bool up = k < SLOTS ? ((s & k) == 0) : ((lane & (k / SLOTS)) == 0);
            const K a = key[s], b = key[s | j];
            const bool sw = (a > b) == up;
            key[s] = sw ? b : a;
            key[s | j] = sw ? a : b;
          }
        }
      } else {
        const int jl = j / SLOTS;                                 // lane distance (k > j >= SLOTS: the direction depends on the lane only)
        const bool take_min = ((lane & (k / SLOTS)) == 0) == ((lane & jl) == 0);
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          const K a = key[s];
          const K o = __shfl_xor_sync(0xffffffffu, a, jl);
          key[s] = ((a < o) == take_min) ? a : o;                 // keeps its own key iff (own < other) == (this place takes the minimum)
        }
      }
    }
  }
}

__host__ __device__ inline int next_pow2_min32(int n) {
#ifdef __CUDA_ARCH__
  return n <= 32 ? 32 : 1 << (32 - __clz(n - 1));
#else
  int p = 32;
  while (p < n) p <<= 1;
  return p;
#endif
}

// 64-bit words of shared memory one warp needs for n samples (S per list): keys[next_pow2(n)], then the fast path's scratch --
// 2 * next_pow2(S) floats of coarse distances (+inf padded) and one occupancy byte per place rounded up to 128 places
__host__ __device__ inline int merge_smem_words(int n, int S) { return next_pow2_min32(n) + next_pow2_min32(S) + S + 16; }

// Fast path for S <= 32 * SLOTS.  t_fine / t_coarse point at the ray's first distance.  Returns false (keys untouched) when the
// coarse distances are not non-decreasing.
template <int SLOTS>
__device__ __forceinline__ bool sort_fine_and_merge(unsigned long long* keys, unsigned long long* aux, const float* __restrict__ t_fine,
                                                    const float* __restrict__ t_coarse, int S, int lane) {
  constexpr int kP2C = 64 * SLOTS;                // > S: the search below never needs a bound check
  constexpr int kIdxBits = SLOTS == 1 ? 5 : (SLOTS == 2 ? 6 : (SLOTS == 4 ? 7 : 8));
  const int n = 2 * S;
  float* ct = reinterpret_cast<float*>(aux);                                   // [kP2C] coarse distances, +inf from S on
  uint32_t* occ = reinterpret_cast<uint32_t*>(ct + kP2C);                      // one byte per merged place, padded to 128 places
  const float inf = __int_as_float(0x7f800000);
  // ---- coarse distances: registers (lane-major) -> sortedness vote -> shared memory ----
  float c[SLOTS];
  bool ok = true;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int e = lane * SLOTS + s;
    c[s] = e < S ? __ldg(t_coarse + e) : inf;
    if (s > 0) ok = ok && (c[s - 1] <= c[s]);
  }
  const float c_next = __shfl_down_sync(0xffffffffu, c[0], 1);
  if (lane < 31) ok = ok && (c[SLOTS - 1] <= c_next);
  if (!__all_sync(0xffffffffu, ok)) return false;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    ct[lane * SLOTS + s] = c[s];
    ct[32 * SLOTS + lane * SLOTS + s] = inf;
  }
  for (int w = lane; w < ((n + 127) >> 7) * 32; w += 32) {                     // places >= n count as occupied
    const int left = n - 4 * w;
    occ[w] = left >= 4 ? 0u : (left <= 0 ? 0x01010101u : (0x01010101u << (8 * left)));
  }
  // ---- fine keys: sort in registers ----
  uint32_t sb[SLOTS];
  uint32_t mn = 0xffffffffu, mx = 0u;
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int e = lane * SLOTS + s;
    sb[s] = 0xffffffffu;
    if (e < S) {
      sb[s] = sortable_bits(__ldg(t_fine + e));
      mn = min(mn, sb[s]);
      mx = max(mx, sb[s]);
    }
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  float tf[SLOTS];
  int src[SLOTS];
  if (mx - mn < (1u << (32 - kIdxBits)) - 1u) {                                // warp-uniform
    uint32_t k32[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int e = lane * SLOTS + s;
      k32[s] = e >= S ? 0xffffffffu : (((sb[s] - mn) << kIdxBits) | static_cast<uint32_t>(e));
    }
    warp_bitonic_sort_lane_major<SLOTS>(k32, lane);
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      src[s] = static_cast<int>(k32[s] & ((1u << kIdxBits) - 1u));
      tf[s] = from_sortable_bits((k32[s] >> kIdxBits) + mn);
    }
  } else {
    unsigned long long k64[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int e = lane * SLOTS + s;
      k64[s] = e >= S ? ~0ull : ((static_cast<unsigned long long>(sb[s]) << 32) | static_cast<unsigned>(e));
    }
    warp_bitonic_sort_lane_major<SLOTS>(k64, lane);
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      src[s] = static_cast<int>(k64[s] & 0xffffffffu);
      tf[s] = from_sortable_bits(static_cast<uint32_t>(k64[s] >> 32));
    }
  }
  __syncwarp();
  // ---- a fine key's place: its rank + the number of coarse distances strictly below it (ties: fine first) ----
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int e = lane * SLOTS + s;                                            // the padding sorted to the end: e >= S
    if (e < S) {
      int r = 0;
#pragma unroll
      for (int step = kP2C >> 1; step > 0; step >>= 1)
        if (ct[r + step - 1] < tf[s]) r += step;
      keys[e + r] = make_key(tf[s], src[s]);
      reinterpret_cast<unsigned char*>(occ)[e + r] = 1;
    }
  }
  __syncwarp();
  // ---- the coarse keys take the empty places in order ----
  int z_carry = 0;
  for (int base = 0; base < n; base += 128) {
    const uint32_t o = occ[(base >> 2) + lane];
    const int z = 4 - __popc(o & 0x01010101u);
    int incl = z;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += up;
    }
    int zi = z_carry + incl - z;                                               // index of this lane's first coarse distance
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (((o >> (8 * i)) & 1u) == 0) {
        keys[base + 4 * lane + i] = make_key(ct[min(zi, S - 1)], S + zi);       // (zi < S unless a NaN broke the ranks)
        ++zi;
      }
    }
    if (n > 128) z_carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  return true;
}

// Generic path: all n keys through the shared-memory network (two lists), or the single list in its given order.
__device__ __forceinline__ void load_and_sort_ray_generic(unsigned long long* keys, const float* __restrict__ t_fine,
                                                          const float* __restrict__ t_coarse, int S, int n, int n2, int lane) {
  const bool two = t_fine != nullptr;
  if (!two) {
    for (int e = lane; e < n; e += 32) keys[e] = make_key(__ldg(t_coarse + e), e);
    __syncwarp();
    return;
  }
  for (int e = lane; e < n2; e += 32) {
    unsigned long long key = ~0ull;
    if (e < n) {
      const float te = e < S ? __ldg(t_fine + e) : __ldg(t_coarse + (e - S));
      key = (static_cast<unsigned long long>(sortable_bits(te)) << 32) | static_cast<unsigned>(e);
    }
    keys[e] = key;
  }
  __syncwarp();
  warp_bitonic_sort(keys, n2, lane);
  for (int e = lane; e < n; e += 32) {
    const unsigned long long k = keys[e];
    keys[e] = make_key(from_sortable_bits(static_cast<uint32_t>(k >> 32)), static_cast<int>(k & 0xffffffffu));
  }
  __syncwarp();
}

// Loads the ray's distances (fine first, then coarse; or coarse only), orders them, and leaves keys[s] = (t, source index).
// kMaxSlots: the largest per-lane key count this instantiation can meet (S <= 32 * kMaxSlots), kMinSlots the smallest worth a
// switch case -- the compositing kernels know the range of S from their samples-per-lane template parameter.
template <int kMinSlots, int kMaxSlots>
__device__ __forceinline__ void load_and_sort_ray(unsigned long long* keys, const float* __restrict__ t_fine,
                                                  const float* __restrict__ t_coarse, long long ray, int S, int n, int n2, int lane) {
  const float* tc = t_coarse + ray * S;
  const float* tf = t_fine ? t_fine + ray * S : nullptr;
  if (tf != nullptr && S <= 32 * kMaxSlots) {
    unsigned long long* aux = keys + n2;
    bool done;
    if (kMinSlots <= 1 && S <= 32) done = sort_fine_and_merge<1>(keys, aux, tf, tc, S, lane);
    else if (kMinSlots <= 2 && kMaxSlots >= 2 && S <= 64) done = sort_fine_and_merge<(kMaxSlots >= 2 ? 2 : 1)>(keys, aux, tf, tc, S, lane);
    else if (kMinSlots <= 4 && kMaxSlots >= 4 && S <= 128) done = sort_fine_and_merge<(kMaxSlots >= 4 ? 4 : 1)>(keys, aux, tf, tc, S, lane);
    else done = sort_fine_and_merge<(kMaxSlots >= 8 ? 8 : 1)>(keys, aux, tf, tc, S, lane);
    if (done) return;
  }
  load_and_sort_ray_generic(keys, tf, tc, S, n, n2, lane);
}

}  // namespace cng
