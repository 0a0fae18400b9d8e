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

__host__ __device__ inline int next_pow2_min32(int n) {
  int p = 32;
  while (p < n) p <<= 1;
  return p;
}

// Loads the ray's distances (fine first, then coarse; or coarse only), sorts, and leaves keys[s] = (t, source index).
__device__ __forceinline__ void load_and_sort_ray(unsigned long long* keys, const float* __restrict__ t_fine,
                                                  const float* __restrict__ t_coarse, long long ray, int S, int n, int n2, int lane) {
  const bool two = t_fine != nullptr;
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
