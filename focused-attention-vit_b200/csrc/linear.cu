// C-ABI entry points of the linear layers: forward, data gradient, weight/bias gradient.
// They replace nn.Linear and its autograd at /root/reference/models/mhla.py:100 (qkv), :158 (proj) and, as the
// "next" row, the MLP linears of models/vit.py:107-139.  bf16 runs on the tcgen05 kernel (gemm_tcgen05.cu), fp32 on
// the SIMT parity kernel (gemm_simt.cu).
#include "favit_common.cuh"
#include "gemm_tcgen05.h"

namespace favit {
namespace {

// db[n] (+)= sum_m dy[m,n].  One thread owns 8 consecutive columns (one 16-byte load per row for bf16), a CTA
// walks a strip of rows, partial sums are combined through shared memory and one atomic per column per CTA.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ dy, float* __restrict__ db, int M, int N,
                                                     int64_t ld, int rows_per_cta) {
  __shared__ float s_part[8][256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta;
  const int r1 = min(M, r0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const bool vec = (c0 + 8 <= N) && (ld % 8 == 0) && ((uintptr_t)dy % 16 == 0);
  if (c0 < N) {
    if (vec) {
      int r = r0 + warp;
      for (; r + 56 < r1; r += 64) {  // 8 independent 16-byte loads in flight per thread
        float f[8][8];
#pragma unroll
        for (int u = 0; u < 8; ++u) load8(dy + (int64_t)(r + 8 * u) * ld + c0, f[u]);
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += f[u][e];
      }
      for (; r < r1; r += 8) {
        float f[8];
        load8(dy + (int64_t)r * ld + c0, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    } else {
      for (int r = r0 + warp; r < r1; r += 8)
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (c0 + e < N) acc[e] += Elem<T>::ld(dy + (int64_t)r * ld + c0 + e);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) s_part[warp][lane * 8 + e] = acc[e];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_part[w][threadIdx.x];
    atomicAdd(db + c, s);
  }
}

template <typename T>
int launch_colsum(const void* dy, float* db, int M, int N, int64_t ld, cudaStream_t st) {
  const int col_blocks = ceil_div(N, 256);
  int row_blocks = max(1, min(ceil_div(M, 64), ceil_div(3 * num_sms(), col_blocks)));
  const int rows_per_cta = ceil_div(ceil_div(M, row_blocks), 8) * 8;
  row_blocks = ceil_div(M, rows_per_cta);
  colsum_kernel<T><<<dim3(col_blocks, row_blocks), 256, 0, st>>>((const T*)dy, db, M, N, ld, rows_per_cta);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

// DropSpec for an [rows, cols] output: p quantised to 16 bits, 4 columns per hash (gemm_epilogue.cuh)
int make_drop(const char* who, float p, const unsigned long long* seed, unsigned long long offset, int cols, DropSpec* d) {
  *d = DropSpec();
  if (p <= 0.f) return FAVIT_OK;
  FAVIT_CHECK_ARG(p < 1.f && seed, "%s: dropout needs 0 <= p < 1 and a device seed pointer (p = %g)", who, (double)p);
  d->thr = (unsigned)lrintf(p * 65536.f);
  if (d->thr == 0) return FAVIT_OK;  // p below 2^-17: nothing is ever dropped
  d->seed = seed;
  d->offset = offset;
  d->inv_keep = 65536.f / (float)(65536u - d->thr);
  d->groups_per_row = (cols + 3) / 4;
  return FAVIT_OK;
}

// out[m, n] = keep(m, n) ? g[m, n] / (1 - p) : 0 in the compute dtype, plus the fp32 column sums of `out` as stored:
// the gradient entering fc2 (models/vit.py:132-138: the dropout after fc2 masks it) as a GEMM operand, and fc2's bias
// gradient.  One thread owns 8 consecutive columns of a strip of rows.
template <typename TOut>
__global__ void __launch_bounds__(256) dropout_cast_kernel(const float* __restrict__ g, TOut* __restrict__ out,
                                                           float* __restrict__ colsum, int M, int N, int rows_per_cta,
                                                           const DropSpec d) {
  __shared__ float s_part[8][256 + 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  const unsigned long long key = d.seed ? __ldg(d.seed) + d.offset : 0ull;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (c0 < N) {   // N % 8 == 0 (checked by the host)
    for (int r = r0 + warp; r < r1; r += 8) {
      float f[8];
      load8(g + (int64_t)r * N + c0, f);
      if (d.seed) {
        const unsigned long long g0 = (unsigned long long)r * (unsigned long long)d.groups_per_row + (unsigned)(c0 >> 2);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const unsigned long long z = splitmix64(key + (g0 + j) * 0x9E3779B97F4A7C15ull);
          const unsigned lo = (unsigned)z, hi = (unsigned)(z >> 32);
          f[4 * j] = (lo & 0xffffu) >= d.thr ? f[4 * j] * d.inv_keep : 0.f;
          f[4 * j + 1] = (lo >> 16) >= d.thr ? f[4 * j + 1] * d.inv_keep : 0.f;
          f[4 * j + 2] = (hi & 0xffffu) >= d.thr ? f[4 * j + 2] * d.inv_keep : 0.f;
          f[4 * j + 3] = (hi >> 16) >= d.thr ? f[4 * j + 3] * d.inv_keep : 0.f;
        }
      }
      store8(out + (int64_t)r * N + c0, f);
      if constexpr (sizeof(TOut) == 2) {   // sum what was stored (bf16-rounded), like the GEMM-epilogue column sums
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __bfloat162float(__float2bfloat16_rn(f[e]));
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  }
  if (!colsum) return;
#pragma unroll
  for (int e = 0; e < 8; ++e) s_part[warp][lane * 8 + e] = acc[e];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_part[w][threadIdx.x];
    atomicAdd(colsum + c, s);
  }
}

int check_dims(const char* who, int M, int N, int K) {
  FAVIT_CHECK_ARG(M > 0 && N > 0 && K > 0, "%s: M,N,K must be positive (got %d,%d,%d)", who, M, N, K);
  return FAVIT_OK;
}

}  // namespace
}  // namespace favit

using namespace favit;

extern "C" int favit_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                                void* preact_out, int M, int N, int K, int64_t ldx, int64_t ldw, int64_t ldy,
                                int64_t ldres, favit_dtype dtype, favit_dtype y_dtype, favit_dtype res_dtype,
                                int epilogue, favit_stream stream) {
  return favit_linear_fwd_dropout(x, w, bias, residual, y, preact_out, M, N, K, ldx, ldw, ldy, ldres, dtype, y_dtype,
                                  res_dtype, epilogue, 0.f, nullptr, 0, stream);
}

extern "C" int favit_linear_fwd_dropout(const void* x, const void* w, const float* bias, const void* residual, void* y,
                                        void* preact_out, int M, int N, int K, int64_t ldx, int64_t ldw, int64_t ldy,
                                        int64_t ldres, favit_dtype dtype, favit_dtype y_dtype, favit_dtype res_dtype,
                                        int epilogue, float drop_p, const uint64_t* drop_seed, uint64_t drop_offset,
                                        favit_stream stream) {
  if (int rc = check_dims("linear_fwd", M, N, K)) return rc;
  FAVIT_CHECK_ARG(x && w && y, "linear_fwd: null x/w/y");
  FAVIT_CHECK_ARG(epilogue == FAVIT_EPI_NONE || epilogue == FAVIT_EPI_GELU, "linear_fwd: bad epilogue %d", epilogue);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FAVIT_F32) {
    FAVIT_CHECK_ARG(drop_p <= 0.f, "linear_fwd: the fused dropout epilogue exists on the bf16 path only");
    FAVIT_CHECK_ARG(y_dtype == FAVIT_F32 && (!residual || res_dtype == FAVIT_F32),
                    "linear_fwd: the fp32 path is fp32 end to end");
    return gemm_simt_launch((const float*)x, (const float*)w, (float*)y, bias, nullptr, (float*)preact_out,
                            (const float*)residual, ldres, M, N, K, ldx, 1, ldw, 1, ldy, epilogue, 1, 0, st);
  }
  FAVIT_CHECK_ARG(dtype == FAVIT_BF16, "linear_fwd: bad dtype");
  tc::Epilogue e;
  e.c = y; e.ldc = ldy; e.c_dtype = y_dtype;
  e.bias = bias;
  e.residual = residual; e.ldres = ldres; e.res_dtype = res_dtype;
  e.aux_out = preact_out; e.ldaux = ldy;
  e.act = epilogue;
  if (int rc = make_drop("linear_fwd", drop_p, (const unsigned long long*)drop_seed, drop_offset, N, &e.drop)) return rc;
  return tc::gemm_bf16(x, 0, ldx, w, 0, ldw, M, N, K, e, 0, 0, st);
}

extern "C" int favit_linear_dgrad(const void* dy, const void* w, const void* preact, void* dx, float* dx_colsum, int M,
                                  int N, int K, int64_t lddy, int64_t ldw, int64_t lddx, favit_dtype dtype,
                                  favit_dtype dx_dtype, int epilogue, favit_stream stream) {
  return favit_linear_dgrad_dropout(dy, w, preact, dx, dx_colsum, M, N, K, lddy, ldw, lddx, dtype, dx_dtype, epilogue,
                                    0.f, nullptr, 0, stream);
}

extern "C" int favit_linear_dgrad_dropout(const void* dy, const void* w, const void* preact, void* dx, float* dx_colsum,
                                          int M, int N, int K, int64_t lddy, int64_t ldw, int64_t lddx, favit_dtype dtype,
                                          favit_dtype dx_dtype, int epilogue, float drop_p, const uint64_t* drop_seed,
                                          uint64_t drop_offset, favit_stream stream) {
  if (int rc = check_dims("linear_dgrad", M, N, K)) return rc;
  FAVIT_CHECK_ARG(dy && w && dx, "linear_dgrad: null dy/w/dx");
  FAVIT_CHECK_ARG(epilogue == FAVIT_EPI_NONE || (epilogue == FAVIT_EPI_DGELU_MUL && preact),
                  "linear_dgrad: bad epilogue %d", epilogue);
  cudaStream_t st = (cudaStream_t)stream;
  // dX[m,k] = sum_n dY[m,n] W[n,k]: output M x K, reduction over N; W is stored [reduction][out] (MN-major B)
  if (dtype == FAVIT_F32) {
    FAVIT_CHECK_ARG(drop_p <= 0.f, "linear_dgrad: the fused dropout epilogue exists on the bf16 path only");
    FAVIT_CHECK_ARG(dx_dtype == FAVIT_F32, "linear_dgrad: the fp32 path is fp32 end to end");
    int rc = gemm_simt_launch((const float*)dy, (const float*)w, (float*)dx, nullptr, (const float*)preact, nullptr,
                              nullptr, 0, M, K, N, lddy, 1, 1, ldw, lddx, epilogue, 1, 0, st);
    if (rc || !dx_colsum) return rc;
    return launch_colsum<float>(dx, dx_colsum, M, K, lddx, st);
  }
  FAVIT_CHECK_ARG(dtype == FAVIT_BF16, "linear_dgrad: bad dtype");
  tc::Epilogue e;
  e.c = dx; e.ldc = lddx; e.c_dtype = dx_dtype;
  e.aux = preact; e.ldaux = lddx;
  e.act = epilogue;
  if (int rc = make_drop("linear_dgrad", drop_p, (const unsigned long long*)drop_seed, drop_offset, K, &e.drop)) return rc;
  // the CTA-pair kernel sums the columns of dX in its epilogue; otherwise one extra pass over dX
  e.colsum = dx_colsum;
  const bool fused = dx_colsum && tc::gemm_bf16_2cta_applicable(M, K, N, e, lddy, ldw);
  if (!fused) e.colsum = nullptr;
  int rc = tc::gemm_bf16(dy, 0, lddy, w, 1, ldw, M, K, N, e, 0, 0, st);
  if (rc || !dx_colsum || fused) return rc;
  if (dx_dtype == FAVIT_BF16) return launch_colsum<__nv_bfloat16>(dx, dx_colsum, M, K, lddx, st);
  return launch_colsum<float>(dx, dx_colsum, M, K, lddx, st);
}

extern "C" int favit_linear_wgrad(const void* dy, const void* x, float* dw, float* db, int M, int N, int K,
                                  int64_t lddy, int64_t ldx, int64_t lddw, favit_dtype dtype, int accumulate,
                                  favit_stream stream) {
  if (int rc = check_dims("linear_wgrad", M, N, K)) return rc;
  FAVIT_CHECK_ARG(dy && x && dw, "linear_wgrad: null dy/x/dw");
  FAVIT_CHECK_ARG(dtype == FAVIT_F32 || dtype == FAVIT_BF16, "linear_wgrad: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    if (lddw == K) {
      FAVIT_CHECK_CUDA(cudaMemsetAsync(dw, 0, (size_t)N * K * sizeof(float), st));
    } else {
      FAVIT_CHECK_CUDA(cudaMemset2DAsync(dw, (size_t)lddw * sizeof(float), 0, (size_t)K * sizeof(float), N, st));
    }
    if (db) FAVIT_CHECK_CUDA(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), st));
  }
  // dW[n,k] = sum_m dY[m,n] X[m,k]: output N x K, reduction over M; both operands stored [reduction][out]
  int rc;
  if (dtype == FAVIT_F32) {
    const int tiles = ceil_div(N, 64) * ceil_div(K, 64);
    int splits = max(1, min(ceil_div(M, 256), (2 * num_sms()) / max(tiles, 1)));
    rc = gemm_simt_launch((const float*)dy, (const float*)x, dw, nullptr, nullptr, nullptr, nullptr, 0, N, K, M, 1,
                          lddy, 1, ldx, lddw, FAVIT_EPI_NONE, splits, 1, st);
  } else {
    tc::Epilogue e;
    e.c = dw; e.ldc = lddw; e.c_dtype = FAVIT_F32;
    e.accumulate = 1; e.split_ok = 1;
    rc = tc::gemm_bf16(dy, 1, lddy, x, 1, ldx, N, K, M, e, 0, 0, st);
  }
  if (rc) return rc;
  if (db) {
    if (dtype == FAVIT_F32) return launch_colsum<float>(dy, db, M, N, lddy, st);
    return launch_colsum<__nv_bfloat16>(dy, db, M, N, lddy, st);
  }
  return FAVIT_OK;
}

extern "C" int favit_gemm_bf16_raw(const void* a, int a_mn, int64_t lda, const void* b, int b_mn, int64_t ldb,
                                   void* c, int64_t ldc, favit_dtype c_dtype, int M, int N, int K, int bn, int splits,
                                   favit_stream stream) {
  if (int rc = check_dims("gemm_bf16_raw", M, N, K)) return rc;
  FAVIT_CHECK_ARG(a && b && c, "gemm_bf16_raw: null pointer");
  tc::Epilogue e;
  e.c = c; e.ldc = ldc; e.c_dtype = c_dtype;
  e.split_ok = (splits != 1) ? 1 : 0;
  return tc::gemm_bf16(a, a_mn ? 1 : 0, lda, b, b_mn ? 1 : 0, ldb, M, N, K, e, bn, splits, (cudaStream_t)stream);
}

extern "C" int favit_set_gemm_tile_scheduler(int mode) { return tc::gemm_tile_scheduler(mode); }

// ---- many tensors per launch: fp32 master weights -> bf16 GEMM operands, small gradients <-> a flat all-reduce buffer ----
namespace favit {
namespace {
constexpr int kCastMax = 32;
struct CastTable {
  const void* src[kCastMax];
  void* dst[kCastMax];
  int64_t n[kCastMax];
};
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) copy_batched_kernel(const __grid_constant__ CastTable t) {
  const TS* s = (const TS*)t.src[blockIdx.y];
  TD* d = (TD*)t.dst[blockIdx.y];
  const int64_t n = t.n[blockIdx.y];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
  const bool vec = ((uintptr_t)s % 16 == 0) && ((uintptr_t)d % 16 == 0);
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (vec && i + 8 <= n) {
      float f[8];
      load8(s + i, f);
      store8(d + i, f);
    } else {
      for (int64_t k = i; k < n && k < i + 8; ++k) Elem<TD>::st(d + k, Elem<TS>::ld(s + k));
    }
  }
}

template <typename TS, typename TD>
int launch_copy_batched(int count, const void* const* src, void* const* dst, const int64_t* numel, cudaStream_t st) {
  for (int t0 = 0; t0 < count; t0 += kCastMax) {
    const int n = count - t0 < kCastMax ? count - t0 : kCastMax;
    CastTable t;
    int64_t big = 0;
    for (int i = 0; i < kCastMax; ++i) {
      const int k = i < n ? t0 + i : t0;  // unused rows repeat the first tensor with length 0
      FAVIT_CHECK_ARG(src[k] && dst[k] && numel[k] >= 0, "copy_batched: null tensor %d", k);
      t.src[i] = src[k];
      t.dst[i] = dst[k];
      t.n[i] = i < n ? numel[k] : 0;
      big = t.n[i] > big ? t.n[i] : big;
    }
    int bx = (int)((big + 2047) / 2048);
    bx = bx < 1 ? 1 : (bx > 4 * num_sms() ? 4 * num_sms() : bx);
    copy_batched_kernel<TS, TD><<<dim3(bx, n), 256, 0, st>>>(t);
    FAVIT_CHECK_LAUNCH();
  }
  return FAVIT_OK;
}
}  // namespace
}  // namespace favit

// dst[t][0..n[t]) = convert(src[t][0..n[t])) for `count` tensors (host arrays of device pointers / sizes), one launch per
// 32 tensors.  Uses: the per-step cast of the fp32 master weights of the MLP (models/vit.py:107-139 fc1 / fc2 of every
// block) into bf16 GEMM operands, and packing the ~100 small gradients of a model (biases, LayerNorm parameters) into one
// flat buffer — and back — so that the data-parallel all-reduce handles one tensor instead of a hundred.
extern "C" int favit_copy_batched(int count, const void* const* src, void* const* dst, const int64_t* numel,
                                  favit_dtype src_dtype, favit_dtype dst_dtype, favit_stream stream) {
  FAVIT_CHECK_ARG(count > 0 && src && dst && numel, "copy_batched: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (src_dtype == FAVIT_F32 && dst_dtype == FAVIT_BF16) return launch_copy_batched<float, __nv_bfloat16>(count, src, dst, numel, st);
  if (src_dtype == FAVIT_BF16 && dst_dtype == FAVIT_F32) return launch_copy_batched<__nv_bfloat16, float>(count, src, dst, numel, st);
  if (src_dtype == FAVIT_F32 && dst_dtype == FAVIT_F32) return launch_copy_batched<float, float>(count, src, dst, numel, st);
  if (src_dtype == FAVIT_BF16 && dst_dtype == FAVIT_BF16)
    return launch_copy_batched<__nv_bfloat16, __nv_bfloat16>(count, src, dst, numel, st);
  set_error("copy_batched: bad dtype");
  return FAVIT_ERR_ARG;
}

extern "C" int favit_cast_bf16_batched(int count, const void* const* src, void* const* dst, const int64_t* numel,
                                       favit_stream stream) {
  return favit_copy_batched(count, src, dst, numel, FAVIT_F32, FAVIT_BF16, stream);
}

// out[n] += sum_m x[m,n]   (accumulates: zero `out` first for a plain column sum)
extern "C" int favit_colsum(const void* x, favit_dtype dtype, float* out, int M, int N, int64_t ld, favit_stream stream) {
  FAVIT_CHECK_ARG(x && out && M > 0 && N > 0, "colsum: bad argument");
  if (dtype == FAVIT_F32) return launch_colsum<float>(x, out, M, N, ld, (cudaStream_t)stream);
  if (dtype == FAVIT_BF16) return launch_colsum<__nv_bfloat16>(x, out, M, N, ld, (cudaStream_t)stream);
  set_error("colsum: bad dtype");
  return FAVIT_ERR_ARG;
}

// out = dropout-masked, rescaled copy of the fp32 gradient g in the compute dtype (+ accumulated fp32 column sums)
extern "C" int favit_dropout_cast(const float* g, void* out, favit_dtype out_dtype, float* colsum, int M, int N,
                                  float drop_p, const uint64_t* drop_seed, uint64_t drop_offset, favit_stream stream) {
  FAVIT_CHECK_ARG(g && out && M > 0 && N > 0, "dropout_cast: bad argument");
  FAVIT_CHECK_ARG(N % 8 == 0 && ((uintptr_t)g % 16 == 0) && ((uintptr_t)out % 16 == 0),
                  "dropout_cast: N must be a multiple of 8 and the tensors 16-byte aligned");
  DropSpec d;
  if (int rc = make_drop("dropout_cast", drop_p, (const unsigned long long*)drop_seed, drop_offset, N, &d)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int col_blocks = ceil_div(N, 256);
  int row_blocks = max(1, min(ceil_div(M, 32), ceil_div(8 * num_sms(), col_blocks)));
  const int rows_per_cta = ceil_div(ceil_div(M, row_blocks), 8) * 8;
  row_blocks = ceil_div(M, rows_per_cta);
  if (out_dtype == FAVIT_BF16)
    dropout_cast_kernel<__nv_bfloat16><<<dim3(col_blocks, row_blocks), 256, 0, st>>>(g, (__nv_bfloat16*)out, colsum, M, N,
                                                                                  rows_per_cta, d);
  else if (out_dtype == FAVIT_F32)
    dropout_cast_kernel<float><<<dim3(col_blocks, row_blocks), 256, 0, st>>>(g, (float*)out, colsum, M, N, rows_per_cta, d);
  else { set_error("dropout_cast: bad dtype"); return FAVIT_ERR_ARG; }
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}
