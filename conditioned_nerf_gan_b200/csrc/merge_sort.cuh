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

__device__ __forceinline__ float key_t_bits(unsigned long long k) { return from_sortable_bits(static_cast<uint32_t>(k >> 32)); }

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
          // keys are unique: the lane keeps its own key iff (own < other) == (this position takes the minimum)
          const unsigned long long a = key[s];
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, a, j);
          const bool take_min = (((s * 32 + lane) & k) == 0) == ((lane & j) == 0);
          key[s] = ((a < o) == take_min) ? a : o;
        }
      }
    }
  }
}

// The network on LANE-MAJOR keys: key index i = lane * SLOTS + slot.  Exchanges at distance < SLOTS stay inside a lane; only the
// 15 stages at lane distance 1..16 of each merge level shuffle (slot-major: every stage below distance 32 does), which is what
// the sort of the fine keys spends most on (a 64-bit shuffle is two SHFL, and SHFL issues at a quarter of the ALU rate).
template <int SLOTS>
__device__ __forceinline__ void warp_bitonic_sort_lane_major(unsigned long long (&key)[SLOTS], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * SLOTS; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j < SLOTS) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
          if ((s & j) == 0) {
            const bool up = k < SLOTS ? ((s & k) == 0) : ((lane & (k / SLOTS)) == 0);
            const unsigned long long a = key[s], b = key[s | j];
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
          const unsigned long long a = key[s];
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, a, jl);
          key[s] = ((a < o) == take_min) ? a : o;
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
// crosses a neighbour): sort only the S fine keys in registers, then MERGE.  A fine key's place in the merged order is its rank
// among the fine keys plus the number of coarse distances strictly below it (ties: fine first, generators.py:163-165 -- the
// stable sort of the fine-first concatenation), found by a branch-free binary search over the coarse distances in shared
// memory; the coarse keys then fill the places the fine keys left empty, in order (an occupancy byte per place, a zero count
// per lane and one warp scan) -- no second search and no shared-memory copy of the sorted fine list.
// scratch after keys[0..n2): `aux` = at least next_pow2_min32(S) + S 64-bit words (merge_smem_words).
template <int SLOTS>
__device__ __forceinline__ void sort_fine_and_merge(unsigned long long* keys, unsigned long long* aux,
                                                    const float* __restrict__ t_fine, const float* __restrict__ t_coarse, long long ray,
                                                    int S, int lane) {
  const int n = 2 * S;
  int p2c = 32;                                   // power of two > S: the search below then never needs a bound check
  while (p2c <= S) p2c <<= 1;
  float* ct = reinterpret_cast<float*>(aux);                                   // [p2c] coarse distances, padded with +inf
  uint32_t* occ = reinterpret_cast<uint32_t*>(ct + p2c);                       // one byte per merged place, padded to 128 places
  const int occ_words = ((n + 127) / 128) * 32;
  unsigned long long key[SLOTS];                  // lane-major: key index e = lane * SLOTS + slot
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int e = lane * SLOTS + s;
    key[s] = ~0ull;
    if (e < S) key[s] = (static_cast<unsigned long long>(sortable_bits(__ldg(t_fine + ray * S + e))) << 32) | static_cast<unsigned>(e);
  }
  for (int e = lane; e < p2c; e += 32) ct[e] = e < S ? __ldg(t_coarse + ray * S + e) : __int_as_float(0x7f800000);
  for (int w = lane; w < occ_words; w += 32) {    // places >= n count as occupied
    const int left = n - 4 * w;
    occ[w] = left >= 4 ? 0u : (left <= 0 ? 0x01010101u : (0x01010101u << (8 * left)));
  }
  warp_bitonic_sort_lane_major<SLOTS>(key, lane);
  __syncwarp();
#pragma unroll
  for (int s = 0; s < SLOTS; ++s) {
    const int e = lane * SLOTS + s;
    if (e < S) {
      const float tf = key_t_bits(key[s]);
      int r = 0;                                                               // number of coarse distances < tf
      for (int step = p2c >> 1; step > 0; step >>= 1)
        if (ct[r + step - 1] < tf) r += step;
      keys[e + r] = key[s];
      reinterpret_cast<unsigned char*>(occ)[e + r] = 1;
    }
  }
  __syncwarp();
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
        keys[base + 4 * lane + i] = (static_cast<unsigned long long>(sortable_bits(ct[min(zi, S - 1)])) << 32) | static_cast<unsigned>(S + zi);   // (zi < S unless a NaN broke the ranks)
        ++zi;
      }
    }
    z_carry += __shfl_sync(0xffffffffu, incl, 31);
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
    unsigned long long* aux = keys + n2;
    switch (next_pow2_min32(S)) {
      case 32: sort_fine_and_merge<1>(keys, aux, t_fine, t_coarse, ray, S, lane); return;
      case 64: sort_fine_and_merge<2>(keys, aux, t_fine, t_coarse, ray, S, lane); return;
      case 128: sort_fine_and_merge<4>(keys, aux, t_fine, t_coarse, ray, S, lane); return;
      default: sort_fine_and_merge<8>(keys, aux, t_fine, t_coarse, ray, S, lane); return;
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
