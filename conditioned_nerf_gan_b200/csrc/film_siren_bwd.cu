// a14, MLP part: the backward of the FiLM-SIREN MLP for one chunk of points of ONE batch item, as a single C-ABI call.
//
// Replaces autograd's walk back through FiLMLayer.forward x L + head (generators/siren.py:146-160, 573-579; what
// loss.backward() runs at utils.py:711) for that chunk:
//   1. recompute: the training-mode tcgen05 forward (film_siren_tc.cu) re-runs the chunk and streams every layer's output
//      x_{l+1} (bf16) and local derivative g_l = freq * cos(u_l) (fp16) to the workspace;
//   2. head: d_o = d_out (* rgb (1 - rgb) on the first three outputs), d_final_w += d_o^T x_L, d_final_b += colsum(d_o),
//      dy = d_o final_w;
//   3. per layer, last to first: dz = dy * g_l (+ column sums; film_grad_from_g_kernel of backward.cu),
//      dW_l += dz^T x_l, dy = dz W_l (d_feat for layer 0); a residual block's kept activation additionally receives the
//      adding layer's dz (res masks as in cng_film_siren_fwd_res).
// The GEMMs are plain library GEMMs (cuBLAS bf16 x bf16 -> fp32 accumulate, cublasGemmEx); cuBLAS is bound at run time
// with dlopen (the library the process already has, e.g. PyTorch's, is reused), so libcng_b200.so has no link-time
// dependency on it.  What the host keeps: the per-item arithmetic that turns the accumulated dW / column sums into
// d_freq, d_phase, d_bias (generators/autograd.py) -- a handful of [L, 256] operations per item.
// Before this entry point existed the same sequence was issued from Python: ~115 launches per chunk through the torch
// dispatcher, which made the train step host-bound at 4 images per GPU.
#include <cublas_v2.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <dlfcn.h>

#include "cng_common.cuh"

namespace cng {

// film_siren_tc.cu
int film_siren_tc_launch(const float* feat, int B, long long N, int C, int HID, int L, const float* const* w, const float* const* b,
                         const float* freq, const float* phase, const float* final_w, const float* final_b_dev, int sigmoid_rgb,
                         int half_operands, void* workspace, size_t workspace_bytes, float* out, void* dump_x, void* dump_g,
                         cudaStream_t stream, unsigned res_save_mask, unsigned res_add_mask, float* res_scratch);
size_t film_siren_tc_workspace(int B, int L);

namespace bwd {

constexpr int kH = 256;

// ---- cuBLAS, bound at run time --------------------------------------------------------------------------------
struct Blas {
  void* lib = nullptr;
  cublasStatus_t (*create)(cublasHandle_t*) = nullptr;
  cublasStatus_t (*set_stream)(cublasHandle_t, cudaStream_t) = nullptr;
  cublasStatus_t (*gemm_ex)(cublasHandle_t, cublasOperation_t, cublasOperation_t, int, int, int, const void*, const void*, cudaDataType,
                            int, const void*, cudaDataType, int, const void*, void*, cudaDataType, int, cublasComputeType_t,
                            cublasGemmAlgo_t) = nullptr;
  cublasHandle_t handle[64] = {};
  bool tried = false;
};
static Blas g_blas;

static int blas_handle(cublasHandle_t* out, cudaStream_t stream) {
  Blas& b = g_blas;
  if (!b.tried) {
    b.tried = true;
    for (const char* name : {"libcublas.so.12", "libcublas.so"}) {
      b.lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD);            // the copy the process already has (PyTorch's), if any
      if (!b.lib) b.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (b.lib) break;
    }
    if (b.lib) {
      b.create = reinterpret_cast<decltype(b.create)>(dlsym(b.lib, "cublasCreate_v2"));
      b.set_stream = reinterpret_cast<decltype(b.set_stream)>(dlsym(b.lib, "cublasSetStream_v2"));
      b.gemm_ex = reinterpret_cast<decltype(b.gemm_ex)>(dlsym(b.lib, "cublasGemmEx"));
    }
  }
  if (!b.lib || !b.create || !b.set_stream || !b.gemm_ex) return fail(CNG_ERR_UNSUPPORTED, "film_siren_bwd: libcublas.so.12 is not loadable (%s)", dlerror());
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fail(CNG_ERR_NO_DEVICE, "film_siren_bwd: no current device");
  if (!b.handle[dev] && b.create(&b.handle[dev]) != CUBLAS_STATUS_SUCCESS) return fail(CNG_ERR_UNSUPPORTED, "film_siren_bwd: cublasCreate failed");
  if (b.set_stream(b.handle[dev], stream) != CUBLAS_STATUS_SUCCESS) return fail(CNG_ERR_UNSUPPORTED, "film_siren_bwd: cublasSetStream failed");
  *out = b.handle[dev];
  return CNG_OK;
}

// Row-major C[M,N] (+)= op(A) op(B) written for cuBLAS' column-major view: C^T = op(B)^T op(A)^T.
//   a_rows_are_k: A is stored [K, M] row-major (i.e. the product uses A^T), else [M, K].  B is always [K, N] row-major.
static int gemm(cublasHandle_t h, bool a_rows_are_k, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C,
                cudaDataType c_type, int ldc, float beta, const char* what) {
  const float alpha = 1.f;
  const cublasStatus_t st = g_blas.gemm_ex(h, CUBLAS_OP_N, a_rows_are_k ? CUBLAS_OP_T : CUBLAS_OP_N, N, M, K, &alpha, B, CUDA_R_16BF, ldb, A,
                                           CUDA_R_16BF, lda, &beta, C, c_type, ldc, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT_TENSOR_OP);
  if (st != CUBLAS_STATUS_SUCCESS) return fail(CNG_ERR_UNSUPPORTED, "film_siren_bwd: cublasGemmEx(%s) failed with status %d", what, static_cast<int>(st));
  return CNG_OK;
}

// ---- small kernels -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 w;
    w.x = *reinterpret_cast<uint32_t*>(&a);
    w.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i) = w;
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
  }
}

constexpr int kHeadRows = 64;
// d_o = d_out (* rgb (1 - rgb) for the colour outputs), d_o_bf16, d_final_b += colsum(d_o), dy = bf16(d_o) . bf16(final_w)
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ out, int sigmoid_rgb,
                                                       const __nv_bfloat16* __restrict__ fw_bf, long long P, __nv_bfloat16* __restrict__ d_o_bf,
                                                       __nv_bfloat16* __restrict__ dy, float* __restrict__ d_fb) {
  __shared__ float s_do[kHeadRows][4];
  const long long r0 = static_cast<long long>(blockIdx.x) * kHeadRows;
  const int rows = static_cast<int>(min(static_cast<long long>(kHeadRows), P - r0));
  const int tid = threadIdx.x;
  {
    const int r = tid >> 2, j = tid & 3;                      // 64 rows x 4 outputs
    float v = 0.f;
    if (r < rows) {
      v = __ldg(d_out + (r0 + r) * 4 + j);
      if (sigmoid_rgb && j < 3) {
        const float y = __ldg(out + (r0 + r) * 4 + j);
        v *= y * (1.f - y);
      }
      d_o_bf[(r0 + r) * 4 + j] = __float2bfloat16_rn(v);
    }
    s_do[r][j] = v;
  }
  __syncthreads();
  if (tid < 4) {
    float a = 0.f;
    for (int r = 0; r < rows; ++r) a += s_do[r][tid];
    atomicAdd(d_fb + tid, a);
  }
  float w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) w[j] = __bfloat162float(fw_bf[j * kH + tid]);
  for (int r = 0; r < rows; ++r) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) a = fmaf(__bfloat162float(__float2bfloat16_rn(s_do[r][j])), w[j], a);
    dy[(r0 + r) * kH + tid] = __float2bfloat16_rn(a);
  }
}

__global__ void __launch_bounds__(256) add_bf16_kernel(__nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) * 8;
  if (i + 7 < n) {
    uint4 x = *reinterpret_cast<const uint4*>(a + i);
    const uint4 y = __ldg(reinterpret_cast<const uint4*>(b + i));
    __nv_bfloat162* xa = reinterpret_cast<__nv_bfloat162*>(&x);
    const __nv_bfloat162* yb = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 p = __bfloat1622float2(xa[k]), q = __bfloat1622float2(yb[k]);
      xa[k] = __floats2bfloat162_rn(p.x + q.x, p.y + q.y);
    }
    *reinterpret_cast<uint4*>(a + i) = x;
  } else {
    for (long long j = i; j < n; ++j) a[j] = __float2bfloat16_rn(__bfloat162float(a[j]) + __bfloat162float(b[j]));
  }
}

struct Layout {
  size_t fold, xs, gs, out, d_o, x0, dy0, dy1, dz[3], total;
};
static Layout layout(long long P, int C, int L) {
  auto up = [](size_t v) { return (v + 255) & ~static_cast<size_t>(255); };
  Layout l{};
  size_t off = 0;
  l.fold = off; off += up(film_siren_tc_workspace(1, L));
  l.xs = off; off += up(static_cast<size_t>(L) * P * kH * 2);
  l.gs = off; off += up(static_cast<size_t>(L) * P * kH * 2);
  l.out = off; off += up(static_cast<size_t>(P) * 16);
  l.d_o = off; off += up(static_cast<size_t>(P) * 8);
  l.x0 = off; off += up(static_cast<size_t>(P) * C * 2);
  l.dy0 = off; off += up(static_cast<size_t>(P) * kH * 2);
  l.dy1 = off; off += up(static_cast<size_t>(P) * kH * 2);
  for (int i = 0; i < 3; ++i) { l.dz[i] = off; off += up(static_cast<size_t>(P) * kH * 2); }
  l.total = off;
  return l;
}

}  // namespace bwd
}  // namespace cng

extern "C" {

size_t cng_film_siren_bwd_workspace_bytes(long long P, int C, int HID, int L) {
  if (P <= 0 || C != 32 || HID != 256 || L < 1 || L > 16) return 0;
  return cng::bwd::layout(P, C, L).total;
}

int cng_film_siren_bwd(const float* feat, const float* d_out, const float* out, long long P, int C, int HID, int L,
                       const float* const* layer_w_host, const float* const* layer_b_host, const void* const* layer_w_bf16_host,
                       const float* freq, const float* phase, const float* final_w, const float* final_b, const void* final_w_bf16,
                       int sigmoid_rgb, unsigned res_save_mask, unsigned res_add_mask, void* workspace, size_t workspace_bytes,
                       void* res_scratch, size_t res_scratch_bytes, float* d_feat, float* const* d_w_acc_host, float* colsum_acc,
                       float* d_final_w_acc, float* d_final_b_acc, cng_stream_t stream) {
  using namespace cng;
  using namespace cng::bwd;
  CNG_REQUIRE(P >= 0 && C == 32 && HID == kH && L >= 1 && L <= 16, CNG_ERR_UNSUPPORTED, "film_siren_bwd: needs C=32, HID=256, L<=16 (got %d, %d, %d)", C, HID, L);
  if (P == 0) return CNG_OK;
  CNG_REQUIRE(P < (1LL << 31), CNG_ERR_UNSUPPORTED, "film_siren_bwd: chunk of %lld points (>= 2^31)", P);
  CNG_REQUIRE(feat && d_out && layer_w_host && layer_b_host && layer_w_bf16_host && freq && phase && final_w && final_b && final_w_bf16 &&
                  d_feat && d_w_acc_host && colsum_acc && d_final_w_acc && d_final_b_acc && workspace,
              CNG_ERR_INVALID_ARGUMENT, "film_siren_bwd: NULL pointer");
  CNG_REQUIRE(!sigmoid_rgb || out, CNG_ERR_INVALID_ARGUMENT, "film_siren_bwd: sigmoid_rgb needs the forward output");
  CNG_REQUIRE(((res_save_mask | res_add_mask) >> L) == 0, CNG_ERR_INVALID_ARGUMENT, "film_siren_bwd: residual mask bit beyond layer %d", L - 1);
  for (int l = 0; l < L; ++l)
    CNG_REQUIRE(layer_w_host[l] && layer_b_host[l] && layer_w_bf16_host[l] && d_w_acc_host[l], CNG_ERR_INVALID_ARGUMENT, "film_siren_bwd: NULL layer %d", l);
  const Layout lay = layout(P, C, L);
  CNG_REQUIRE(workspace_bytes >= lay.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, CNG_ERR_WORKSPACE,
              "film_siren_bwd: workspace %zu < %zu bytes (or not 256-byte aligned)", workspace_bytes, lay.total);
  const bool res = (res_save_mask | res_add_mask) != 0;
  CNG_REQUIRE(!res || (res_scratch && res_scratch_bytes >= cng_film_siren_res_scratch_bytes()), CNG_ERR_WORKSPACE,
              "film_siren_bwd: residual scratch %zu < %zu bytes", res_scratch_bytes, cng_film_siren_res_scratch_bytes());
  if (int e = cng_device_check()) return e;
  cudaStream_t st = as_stream(stream);
  cublasHandle_t h;
  if (int e = blas_handle(&h, st)) return e;

  uint8_t* ws = static_cast<uint8_t*>(workspace);
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(ws + lay.xs);
  __half* gs = reinterpret_cast<__half*>(ws + lay.gs);
  float* out_tmp = reinterpret_cast<float*>(ws + lay.out);
  __nv_bfloat16* d_o_bf = reinterpret_cast<__nv_bfloat16*>(ws + lay.d_o);
  __nv_bfloat16* x0_bf = reinterpret_cast<__nv_bfloat16*>(ws + lay.x0);
  __nv_bfloat16* dy_buf[2] = {reinterpret_cast<__nv_bfloat16*>(ws + lay.dy0), reinterpret_cast<__nv_bfloat16*>(ws + lay.dy1)};
  __nv_bfloat16* dz_buf[3] = {reinterpret_cast<__nv_bfloat16*>(ws + lay.dz[0]), reinterpret_cast<__nv_bfloat16*>(ws + lay.dz[1]),
                              reinterpret_cast<__nv_bfloat16*>(ws + lay.dz[2])};
  const size_t per_layer = static_cast<size_t>(P) * kH;

  // 1. recompute with dumps (one item: B = 1)
  if (int e = film_siren_tc_launch(feat, 1, P, C, HID, L, layer_w_host, layer_b_host, freq, phase, final_w, final_b, sigmoid_rgb, 0, ws + lay.fold,
                                   film_siren_tc_workspace(1, L), out_tmp, xs, gs, st, res_save_mask, res_add_mask, static_cast<float*>(res_scratch)))
    return e;
  // 2. head
  const unsigned head_blocks = static_cast<unsigned>((P + kHeadRows - 1) / kHeadRows);
  head_bwd_kernel<<<head_blocks, 256, 0, st>>>(d_out, out, sigmoid_rgb, static_cast<const __nv_bfloat16*>(final_w_bf16), P, d_o_bf, dy_buf[0], d_final_b_acc);
  if (int e = check_launch("cng_film_siren_bwd: head")) return e;
  // d_final_w[4, H] += d_o^T[4, P] x_L[P, H]
  if (int e = gemm(h, true, 4, kH, static_cast<int>(P), d_o_bf, 4, xs + static_cast<size_t>(L - 1) * per_layer, kH, d_final_w_acc, CUDA_R_32F, kH, 1.f, "d_final_w"))
    return e;
  {
    const long long n = P * C;
    to_bf16_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256 + 1), 256, 0, st>>>(feat, x0_bf, n);
    if (int e = check_launch("cng_film_siren_bwd: features to bf16")) return e;
  }
  // 3. layers, last to first
  int cur = 0;                                                   // dy_buf[cur] holds dy of the layer being processed
  const __nv_bfloat16* pending[16] = {};                        // save layer -> dz arriving through the skip
  for (int l = L - 1; l >= 0; --l) {
    __nv_bfloat16* dz = dz_buf[l % 3];
    if (int e = cng_film_grad_from_g(dy_buf[cur], gs + static_cast<size_t>(l) * per_layer, P, kH, dz, colsum_acc + static_cast<size_t>(l) * kH, stream)) return e;
    if ((res_add_mask >> l) & 1u) {
      int kept = -1;
      for (int s = 0; s < l; ++s)
        if ((res_save_mask >> s) & 1u) kept = s;
      CNG_REQUIRE(kept >= 0, CNG_ERR_INVALID_ARGUMENT, "film_siren_bwd: layer %d adds a residual that no earlier layer kept", l);
      pending[kept] = dz;
    }
    const int K = (l == 0) ? C : kH;
    const __nv_bfloat16* x_in = (l == 0) ? x0_bf : xs + static_cast<size_t>(l - 1) * per_layer;
    // dW_l[H, K] += dz^T[H, P] x_in[P, K]
    if (int e = gemm(h, true, kH, K, static_cast<int>(P), dz, kH, x_in, K, d_w_acc_host[l], CUDA_R_32F, K, 1.f, "dW")) return e;
    if (l == 0) {
      // d_feat[P, C] = dz[P, H] W_0[H, C]
      if (int e = gemm(h, false, static_cast<int>(P), C, kH, dz, kH, layer_w_bf16_host[0], C, d_feat, CUDA_R_32F, C, 0.f, "d_feat")) return e;
    } else {
      __nv_bfloat16* dy_next = dy_buf[cur ^ 1];
      if (int e = gemm(h, false, static_cast<int>(P), kH, kH, dz, kH, layer_w_bf16_host[l], kH, dy_next, CUDA_R_16BF, kH, 0.f, "dy")) return e;
      if (pending[l - 1] != nullptr) {
        const long long n = P * kH;
        add_bf16_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256 + 1), 256, 0, st>>>(dy_next, pending[l - 1], n);
        if (int e = check_launch("cng_film_siren_bwd: residual gradient")) return e;
        pending[l - 1] = nullptr;
      }
      cur ^= 1;
    }
  }
  return CNG_OK;
}

}  // extern "C"
