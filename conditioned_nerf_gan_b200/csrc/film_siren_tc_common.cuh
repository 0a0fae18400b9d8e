// Pieces shared by the two tcgen05 FiLM-SIREN kernels (film_siren_tc.cu: one CTA per SM, cta_group::1;
// film_siren_tc2.cu: CTA pairs, cta_group::2): tile constants, the weight-image layout written by the fold
// kernel, PTX wrappers (mbarrier, bulk copy, tcgen05 fences / ld / st / commit), descriptors.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "cng_common.cuh"

#ifndef CNG_MBAR_HINT_NS
#define CNG_MBAR_HINT_NS 20000     // suspend-time hint of mbarrier.try_wait (A/B knob)
#endif

namespace cng {

constexpr int kHID = 256;
constexpr int kC0 = 32;
constexpr int kTileM = 128;
constexpr int kChunkBytes = 32768;          // [256 n][64 k] bf16
constexpr int kHeadBytes = 8192;            // 4 x [16 n][64 k] bf16
constexpr int kABlockBytes = 16384;         // [128 m][64 k] bf16
constexpr int kATileBytes = 4 * kABlockBytes;

// ---- workspace layout ------------------------------------------------------------------------
// per item: [L0c0][L0c1][L1c0..L1c3]...[L(L-1)c3][head]  then, after all items, shift[B][L][256]
__host__ __device__ inline size_t item_image_bytes(int L) {
  return static_cast<size_t>(2 + 4 * (L - 1)) * kChunkBytes + kHeadBytes;
}
__host__ __device__ inline size_t chunk_offset(int L, int l, int c) {     // l == L -> head
  if (l == 0) return static_cast<size_t>(c) * kChunkBytes;
  if (l < L) return static_cast<size_t>(2 + 4 * (l - 1) + c) * kChunkBytes;
  return static_cast<size_t>(2 + 4 * (L - 1)) * kChunkBytes;
}

// byte offset of bf16 element (row, k) inside a [rows][64] K-major SWIZZLE_128B block
__host__ __device__ inline uint32_t sw128_offset(int row, int k) {
  return static_cast<uint32_t>(row) * 128u + ((((static_cast<uint32_t>(k) >> 3) ^ (row & 7)) << 4)) + (k & 7) * 2u;
}

// ---- PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait suspends the thread in hardware (up to the hint) instead of spinning through issue slots
// that the other tile slot's epilogue warps need; a protocol bug turns into a trap (launch failure)
// after ~4 s instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(static_cast<uint32_t>(CNG_MBAR_HINT_NS))
        : "memory");
    if (ok) break;
    if ((it & 63u) == 63u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) __trap();
    }
  }
}
// Lean mbarrier wait for the MMA-issuing warp (its loop is instruction-latency bound: every SASS instruction between
// two tcgen05.mma issues is ~4 cycles of exposed latency).  Fast path = one try_wait + one branch; a protocol bug
// still ends in a trap instead of a hung GPU.
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DONE;\n\t"
      "add.u32 n, n, 1;\n\t"
      "setp.lt.u32 p, n, 4000000;\n\t"
      "@p bra LAB_WAIT;\n\t"
      "trap;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(static_cast<uint32_t>(CNG_MBAR_HINT_NS))
      : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// L2 eviction priorities for bulk copies and streaming stores (the 64-bit policy words createpolicy.fractional.L2::evict_* 1.0
// produces; the same constants CUTLASS passes as TMA cache hints).  The training-mode kernels stream gigabytes of dumps through
// L2 while every CTA re-reads the same ~1 MB of weight images: without priorities the stream evicts the weights and the weight
// ring starves the tensor pipe (measured: 5000 instead of 2650 cycles per tile-layer of MMA in the training forward).
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst, uint32_t src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void st_global_evict_first(uint4* dst, const uint4& v) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(kL2EvictFirst)
               : "memory");
}
__device__ __forceinline__ uint4 ld_global_evict_first(const uint4* src) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.b32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src), "l"(kL2EvictFirst));
  return v;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared -> global bulk copies (TMA unit, bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// true in exactly one lane of the (converged) warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 in
// [0,14), LBO>>4 in [16,30) (=1, unused for swizzled K-major), SBO>>4 in [32,46) (8 rows x 128 B =
// 1024), version=1 at bit 46, layout_type=2 (SWIZZLE_128B) at [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: c=F32 (1<<4), a=b=BF16 (1<<7, 1<<10), K-major both, N>>3 at 17, M>>4 at 24
__device__ __forceinline__ constexpr uint32_t make_idesc(int M, int N, bool half_operands = false) {
  // a_format / b_format: 0 = F16, 1 = BF16 (kind::f16 serves both at the same rate)
  const uint32_t fmt = half_operands ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

#define CNG_TMEM_LD_32(taddr, v)                                                                                      \
  asm volatile(                                                                                                       \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                       \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                       \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                       \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),        \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),       \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                     \
      : "r"(taddr)                                                                                                    \
      : "memory")

#define CNG_TMEM_LD_16(taddr, v)                                                                                      \
  asm volatile(                                                                                                       \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                       \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                                \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),   \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                      \
      : "r"(taddr)                                                                                                    \
      : "memory")

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 16 consecutive fp32 columns of this warp's 32 lanes <- the same 16 values in every lane
__device__ __forceinline__ void tmem_st_16(uint32_t taddr, const float4& a, const float4& b, const float4& c, const float4& d) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w), "f"(c.x), "f"(c.y),
        "f"(c.z), "f"(c.w), "f"(d.x), "f"(d.y), "f"(d.z), "f"(d.w)
      : "memory");
}
// 32 shift values (warp-uniform address) held in registers, loaded one block ahead of their use so the
// L2 latency of the load is hidden behind the sines of the previous block
struct Shift32 {
  float4 v[8];
  __device__ __forceinline__ void load(const float* __restrict__ shift) {
    const float4* s4 = reinterpret_cast<const float4*>(shift);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __ldg(s4 + i);
  }
  // 32 accumulator columns starting at `taddr` <- the 32 values, identical in every lane
  __device__ __forceinline__ void store(uint32_t taddr) const {
    tmem_st_16(taddr, v[0], v[1], v[2], v[3]);
    tmem_st_16(taddr + 16, v[4], v[5], v[6], v[7]);
  }
};

// the same at 16-column granularity (software-pipelined epilogue)
struct Shift16 {
  float4 v[4];
  __device__ __forceinline__ void load(const float* __restrict__ shift) {
    const float4* s4 = reinterpret_cast<const float4*>(shift);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __ldg(s4 + i);
  }
  __device__ __forceinline__ void store(uint32_t taddr) const { tmem_st_16(taddr, v[0], v[1], v[2], v[3]); }
};

// sin(x) on the FMA/ALU pipes, for the share of elements taken off the MUFU unit: u = x/pi, k = rint(u)
// (magic-number rounding), f = u - k in [-0.5, 0.5], sin(x) = (-1)^k sin(pi f) with an odd degree-5 minimax
// polynomial (max error 6.8e-5, below half a bf16 ulp of the result it feeds).
__device__ __forceinline__ float sin_fma(float x) {
  const float kMagic = 12582912.f;                      // 1.5 * 2^23
  const float t = fmaf(x, 0.31830988618379067f, kMagic);
  const float k = t - kMagic;
  const float f = fmaf(x, 0.31830988618379067f, -k);
  const float f2 = f * f;
  float p = fmaf(f2, 2.2995474338531494f, -5.136905193328857f);
  p = fmaf(p, f2, 3.1406400203704834f);
  const uint32_t sign = __float_as_uint(t) << 31;       // parity of k
  return __uint_as_float(__float_as_uint(p * f) ^ sign);
}
// sin(x) and cos(x) together on the FMA/ALU pipes (shared range reduction; cos(pi f) as an even degree-6 minimax polynomial,
// max error 6.7e-6): the training-mode epilogue needs both per element and would otherwise be MUFU-bound (2 ops per element)
__device__ __forceinline__ void sincos_fma(float x, float& s, float& c) {
  const float kMagic = 12582912.f;
  const float t = fmaf(x, 0.31830988618379067f, kMagic);
  const float k = t - kMagic;
  const float f = fmaf(x, 0.31830988618379067f, -k);
  const float f2 = f * f;
  float ps = fmaf(f2, 2.2995474338531494f, -5.136905193328857f);
  ps = fmaf(ps, f2, 3.1406400203704834f);
  float pc = fmaf(f2, -1.222126841545105f, 4.04128360748291f);
  pc = fmaf(pc, f2, -4.933938026428223f);
  pc = fmaf(pc, f2, 0.9999933242797852f);
  const uint32_t sign = __float_as_uint(t) << 31;
  s = __uint_as_float(__float_as_uint(ps * f) ^ sign);
  c = __uint_as_float(__float_as_uint(pc) ^ sign);
}
template <int kPolyOneIn>
__device__ __forceinline__ void film_sincos(float x, int j, float& s, float& c) {
  if (kPolyOneIn > 0 && (j % (kPolyOneIn > 0 ? kPolyOneIn : 1)) == kPolyOneIn - 1) sincos_fma(x, s, c);
  else __sincosf(x, &s, &c);
}
template <int kPolyOneIn>
__device__ __forceinline__ float film_sin(float x, int j) {
  if (kPolyOneIn > 0 && (j % (kPolyOneIn > 0 ? kPolyOneIn : 1)) == kPolyOneIn - 1) return sin_fma(x);
  return __sinf(x);
}

// 16-bit operand format of the tensor-core path: bf16 (8-bit significand) or fp16 (11-bit; what the reference's
// own autocast uses -- same tcgen05 rate, 8x smaller operand rounding, values here are all well inside its range)
template <bool kHalf>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  if (kHalf) asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));     // first source -> upper half
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <bool kHalf>
__device__ __forceinline__ uint16_t to16(float v) {
  return kHalf ? __half_as_ushort(__float2half_rn(v)) : __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
template <bool kHalf>
__device__ __forceinline__ float from16(uint16_t h) {
  return kHalf ? __half2float(__ushort_as_half(h)) : __bfloat162float(__ushort_as_bfloat16(h));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));   // first source -> upper half
  return r;
}

struct TcParams {
  const float* feat;        // [B, N, 32]; NULL in fused-gather mode:
  // fused gather (kGather instantiations): the prologue of a tile looks its 128 points up in the NDHWC volume itself
  // (trilinear, ATen index arithmetic: trilinear.cuh) instead of reading gathered features back from HBM
  const float4* vol;        // [B, D, H, W, 8] float4
  long long vol_item_stride;   // float4 units; 0 = every item reads item 0's volume
  int D, H, W;
  const float* points;      // [B, N, 3] world positions
  long long N;
  int B, L;
  const uint8_t* images;    // fold output
  const float* shift;       // [B][L][256]
  const float* final_b;     // [4]
  int sigmoid_rgb;
  float* out;               // [B, N, 4]
  long long tiles_per_item;
  long long total_tiles;
  int half_operands;        // 1: fp16 operands, 0: bf16
  // training-mode dumps (backward recompute), NULL in inference.  T = total_tiles:
  //   dump_x  [L][T][64 KB]  the layer output x_{l+1} = sin(u_l) as the 128-point operand TILE IMAGE the next layer's MMA reads
  //           ([4 K-blocks][128 rows][64 x 16 bit], 128B-swizzled, operand format) -- one bulk store per tile-layer; the
  //           weight-gradient kernel (film_siren_bwd_tc.cu) reads it back with MN-major descriptors
  //   dump_g  [L][T][64 KB | 32 KB]  the local derivative g_l = cos(u_l) (the FiLM frequency is folded into the backward's weight
  //           images instead) in the epilogue's own register order -- every store / load is 512 contiguous bytes per warp -- as
  //           fp16, [32-column block cc][lane quarter q][16-byte piece i][lane], or, with g_bits == 8, as codes round(127 cos) + 128,
  //           [cc][q][16-column half h][lane] x 16 B (cng_film_siren_g_dump_bits)
  //   dump_feat [T][16 KB]   the layer-0 operand block [x_hi(32) | x_lo(32)]
  uint8_t* dump_x;
  uint8_t* dump_g;
  uint8_t* dump_feat;
  int g_bits;               // format of dump_g: 16 = fp16 (64 KB per tile-layer, [cc][q][i][lane] x 16 B), 8 = codes (32 KB, [cc][q][h][lane] x 16 B)
  long long* trace;         // debug: clock64 timeline of CTA 0 (tools/trace_tc.py), NULL in production
  // residual blocks (TALLSIREN_dRes, generators/siren.py:218-230): bit l of res_save_mask = layer l's output is kept as the
  // residual, bit l of res_add_mask = the kept residual is added to layer l's pre-activation.  res_scratch: per CTA and
  // tile slot a [64 column quads][128 rows] float4 block (128 KB, stays in L2), lane-contiguous so every access coalesces.
  uint32_t res_save_mask, res_add_mask;
  float* res_scratch;
};
constexpr size_t kResScratchPerSlot = static_cast<size_t>(kHID / 4) * kTileM * sizeof(float4);      // 131072
// trace layout: [iter < 4][layer <= 8][slot < 2][event < 8]; events: 0 MMA thread saw act_ready, 1 MMAs issued,
// 2 epilogue (warp 0 of the slot) saw acc_full, 3 epilogue done (before its arrive), 4 cycles the MMA thread
// spent waiting for weight blocks of this slot-layer, 5 cycles it spent issuing MMAs + commits
#ifndef CNG_TC_TRACE
#define CNG_TC_TRACE 1        // 0: compile the clock64 timeline hooks out
#endif
__device__ __forceinline__ void trace_event(long long* trace, int iter, int l, int x, int ev) {
  if (CNG_TC_TRACE && trace != nullptr && blockIdx.x == 0 && iter < 3 && l <= 8) trace[((iter * 9 + l) * 2 + x) * 8 + ev] = clock64();
}
__device__ __forceinline__ void trace_value(long long* trace, int iter, int l, int x, int ev, long long v) {
  if (CNG_TC_TRACE && trace != nullptr && blockIdx.x == 0 && iter < 3 && l <= 8) trace[((iter * 9 + l) * 2 + x) * 8 + ev] = v;
}

struct TileInfo {
  int item;
  long long n0;
  int rows;
};
__device__ __forceinline__ TileInfo tile_info(const TcParams& p, long long t) {
  TileInfo ti;
  ti.item = static_cast<int>(t / p.tiles_per_item);
  ti.n0 = (t - static_cast<long long>(ti.item) * p.tiles_per_item) * kTileM;
  ti.rows = static_cast<int>(min(static_cast<long long>(kTileM), p.N - ti.n0));
  return ti;
}

// format of the cos(u) dump shared by the training-mode forward and the dgrad chain (16 or 8 bits per element): film_siren_tc.cu
int film_siren_g_dump_bits();

// kPolyOneIn: 0 = every sine on the MUFU unit; n > 0 = one element in n uses sin_fma instead

}  // namespace cng
