// LayerNorm forward / backward for the transformer block around MHLA (sm_100a).
//
// "Next" row of SURVEY.md §8f: the block wrapper /root/reference/models/vit_mhla.py:77-109 (norm1 -> attn -> residual,
// norm2 -> mlp -> residual) runs nn.LayerNorm in fp32 under autocast and hands bf16 to the linears.  These kernels keep
// that numerics (fp32 statistics, two-pass variance) and fuse what surrounds the norm:
//   forward : y = (x - mean) * rstd * gamma + beta, written directly in the GEMM's operand dtype (bf16), mean/rstd saved
//   backward: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
//             + the residual-stream gradient that bypasses the norm (dres), written as fp32 and, optionally, a bf16 copy
//             that the next dgrad / wgrad GEMM reads as its operand; dgamma / dbeta are accumulated per column.
// HBM-bound byte movers: one warp per row, 16-byte vector loads, the row stays in registers between the passes.
#include "favit_common.cuh"

namespace favit {
namespace {

constexpr int kMaxVec = 8;  // float4 per lane: rows up to 32 * 4 * 8 = 1024 elements live in registers

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&f)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&f)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&f)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&f)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&f)[4]) {
  uint2 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

// D % 4 == 0, D <= 1024.  Lane l owns float4 chunks l, l+32, ... (coalesced 512-byte warp accesses).
template <typename TX, typename TY, int NV>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TX* __restrict__ x, const TY* __restrict__ delta,
                                                     TX* __restrict__ xsum, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, TY* __restrict__ y,
                                                     float* __restrict__ mean, float* __restrict__ rstd, int M, int D,
                                                     float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nvec = D >> 2;
  const TX* xr = x + (int64_t)row * D;
  float v[NV][4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) load4(xr + 4 * c, v[i]);
  }
  if (delta) {  // residual add fused in front of the norm: xsum = x + delta is the new residual stream
    float d[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) load4(delta + (int64_t)row * D + 4 * c, d[i]);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[i][e] += d[i][e];
        store4(xsum + (int64_t)row * D + 4 * c, v[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
  }
  const float mu = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = v[i][e] - mu;
        q = fmaf(d, d, q);
      }
    }
  }
  const float rs = rsqrtf(warp_sum(q) / (float)D + eps);
  TY* yr = y + (int64_t)row * D;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      float g[4], b[4], o[4];
      load4(gamma + 4 * c, g);
      load4(beta + 4 * c, b);
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = fmaf((v[i][e] - mu) * rs, g[e], b[e]);
      store4(yr + 4 * c, o);
    }
  }
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// Two warps per row (each owns one half of the columns, NVH float4 per lane), two rows in flight per 128-thread CTA,
// persistent over rows.  The operands of a row (x, dy, dres) are staged through shared memory with cp.async, two rows
// deep: an in-flight cp.async holds no registers, so a CTA keeps ~2 x 7.7 KB per row pair on the wire while it does
// the reductions of the current row (ncu, round 1: the register-staged version sat at 20 % warp occupancy with 10
// long-scoreboard stalls per issue and 3.6 TB/s).  Every thread reads back exactly the 16-byte slots it copied, so the
// staging needs no barrier.  The two row sums are exchanged between the warps of a pair through shared memory and a
// 64-thread named barrier (parity double-buffered).
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
template <typename T> __device__ __forceinline__ void cp_async_vec4(void* smem, const T* gmem);
template <> __device__ __forceinline__ void cp_async_vec4<float>(void* smem, const float* gmem) { cp_async_16(smem, gmem); }
template <> __device__ __forceinline__ void cp_async_vec4<__nv_bfloat16>(void* smem, const __nv_bfloat16* gmem) {
  cp_async_8(smem, gmem);
}

template <typename TX, typename TDY, int NVH>
__global__ void __launch_bounds__(128) ln_bwd_kernel(const TDY* __restrict__ dy, const TX* __restrict__ x,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ dres,
                                                     float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_bf16,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     float* __restrict__ dxsum, int M, int D) {
  // staging: [stage 2][kind 3 = x, dy, dres][NVH][128 threads] 16-byte slots; reused as [3][pairs][D] floats at the end
  extern __shared__ __align__(16) uint8_t s_dyn[];
  __shared__ float2 s_xchg[2][2][2];        // [parity][pair][half] = (sum g, sum g*xhat)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pair = warp >> 1, half = warp & 1, npairs = blockDim.x >> 6;
  const int nvec = D >> 2;
  const int nvh = (nvec + 1) >> 1;           // float4 chunks per half
  const int cbase = half * nvh;
  const int cend = min(nvec, cbase + nvh);
  auto slot = [&](int stage, int kind, int i) -> uint8_t* {
    return s_dyn + ((size_t)((stage * 3 + kind) * NVH + i) * 128 + tid) * 16;
  };
  auto prefetch = [&](int row, int stage) {
    if (row < M) {
#pragma unroll
      for (int i = 0; i < NVH; ++i) {
        const int c = cbase + lane + 32 * i;
        if (c < cend) {
          cp_async_vec4<TX>(slot(stage, 0, i), x + (int64_t)row * D + 4 * c);
          cp_async_vec4<TDY>(slot(stage, 1, i), dy + (int64_t)row * D + 4 * c);
          if (dres) cp_async_16(slot(stage, 2, i), dres + (int64_t)row * D + 4 * c);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float dg[NVH][4], db[NVH][4], ds[NVH][4];   // ds: column sums of dx = the bias gradient of the linear that fed x
#pragma unroll
  for (int i = 0; i < NVH; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dg[i][e] = 0.f; db[i][e] = 0.f; ds[i][e] = 0.f; }
  const float invD = 1.f / (float)D;
  const int stride = gridDim.x * npairs;
  int row = blockIdx.x * npairs + pair;
  int parity = 0;
  prefetch(row, 0);
  float mu_next = row < M ? mean[row] : 0.f, rs_next = row < M ? rstd[row] : 0.f;
  for (; row < M; row += stride, parity ^= 1) {
    prefetch(row + stride, parity ^ 1);
    const float mu = mu_next, rs = rs_next;
    if (row + stride < M) {   // the next row's statistics travel with its operands, not after them
      mu_next = mean[row + stride];
      rs_next = rstd[row + stride];
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // this row's copies (issued by this thread) have landed
    float xh[NVH][4], gy[NVH][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NVH; ++i) {
      const int c = cbase + lane + 32 * i;
      if (c < cend) {
        float g[4];
        load4(reinterpret_cast<const TX*>(slot(parity, 0, i)), xh[i]);
        load4(reinterpret_cast<const TDY*>(slot(parity, 1, i)), gy[i]);
        load4(gamma + 4 * c, g);  // L1-resident
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xv = (xh[i][e] - mu) * rs;
          const float dv = gy[i][e];
          xh[i][e] = xv;
          gy[i][e] = dv * g[e];
          s1 += gy[i][e];
          s2 = fmaf(gy[i][e], xv, s2);
          dg[i][e] = fmaf(dv, xv, dg[i][e]);
          db[i][e] += dv;
        }
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) s_xchg[parity][pair][half] = make_float2(s1, s2);
    asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
    const float2 other = s_xchg[parity][pair][half ^ 1];
    s1 = (s1 + other.x) * invD;
    s2 = (s2 + other.y) * invD;
#pragma unroll
    for (int i = 0; i < NVH; ++i) {
      const int c = cbase + lane + 32 * i;
      if (c < cend) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = rs * (gy[i][e] - s1 - xh[i][e] * s2);
        if (dres) {
          float r[4];
          load4(reinterpret_cast<const float*>(slot(parity, 2, i)), r);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] += r[e];
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) ds[i][e] += o[e];
        store4(dx + (int64_t)row * D + 4 * c, o);
        if (dx_bf16) store4(dx_bf16 + (int64_t)row * D + 4 * c, o);
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (dgamma == nullptr) return;
  __syncthreads();  // the staging area becomes the reduction scratch
  float* sg = reinterpret_cast<float*>(s_dyn);
  float* sb = sg + (size_t)npairs * D;
  float* ss = sg + (size_t)2 * npairs * D;
#pragma unroll
  for (int i = 0; i < NVH; ++i) {
    const int c = cbase + lane + 32 * i;
    if (c < cend) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        sg[pair * D + 4 * c + e] = dg[i][e];
        sb[pair * D + 4 * c + e] = db[i][e];
        ss[pair * D + 4 * c + e] = ds[i][e];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.f, b = 0.f, d = 0.f;
    for (int w = 0; w < npairs; ++w) {
      a += sg[w * D + c];
      b += sb[w * D + c];
      d += ss[w * D + c];
    }
    atomicAdd(dgamma + c, a);
    atomicAdd(dbeta + c, b);
    if (dxsum) atomicAdd(dxsum + c, d);
  }
}

}  // namespace
}  // namespace favit

using namespace favit;

extern "C" int favit_layernorm_fwd(const void* x, favit_dtype x_dtype, const void* delta, void* xsum,
                                   const float* gamma, const float* beta, void* y, favit_dtype y_dtype, float* mean,
                                   float* rstd, int M, int D, float eps, favit_stream stream) {
  FAVIT_CHECK_ARG(x && gamma && beta && y && mean && rstd, "layernorm_fwd: null pointer");
  FAVIT_CHECK_ARG((delta == nullptr) == (xsum == nullptr), "layernorm_fwd: delta and xsum come together");
  FAVIT_CHECK_ARG(M > 0 && D > 0, "layernorm_fwd: M, D must be positive");
  if (D % 4 != 0 || D > 128 * kMaxVec) {
    set_error("layernorm_fwd: D=%d unsupported (needs D %% 4 == 0 and D <= %d)", D, 128 * kMaxVec);
    return FAVIT_ERR_UNSUPPORTED;
  }
  FAVIT_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)gamma % 16 == 0) &&
                      ((uintptr_t)beta % 16 == 0),
                  "layernorm_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)ceil_div(M, 8);
#define LN_FWD_NV(TX, TY, NV) \
  ln_fwd_kernel<TX, TY, NV><<<blocks, 256, 0, st>>>((const TX*)x, (const TY*)delta, (TX*)xsum, gamma, beta, (TY*)y, \
                                                    mean, rstd, M, D, eps)
#define LN_FWD(TX, TY)                            \
  do {                                            \
    if (D <= 256) LN_FWD_NV(TX, TY, 2);           \
    else if (D <= 512) LN_FWD_NV(TX, TY, 4);      \
    else if (D <= 768) LN_FWD_NV(TX, TY, 6);      \
    else LN_FWD_NV(TX, TY, 8);                    \
  } while (0)
  if (x_dtype == FAVIT_F32 && y_dtype == FAVIT_BF16) LN_FWD(float, __nv_bfloat16);
  else if (x_dtype == FAVIT_F32 && y_dtype == FAVIT_F32) LN_FWD(float, float);
  else if (x_dtype == FAVIT_BF16 && y_dtype == FAVIT_BF16) LN_FWD(__nv_bfloat16, __nv_bfloat16);
  else if (x_dtype == FAVIT_BF16 && y_dtype == FAVIT_F32) LN_FWD(__nv_bfloat16, float);
  else {
    set_error("layernorm_fwd: bad dtype");
    return FAVIT_ERR_ARG;
  }
#undef LN_FWD
#undef LN_FWD_NV
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

extern "C" int favit_layernorm_bwd(const void* dy, favit_dtype dy_dtype, const void* x, favit_dtype x_dtype,
                                   const float* mean, const float* rstd, const float* gamma, const float* dres,
                                   float* dx, void* dx_bf16, float* dgamma, float* dbeta, float* dxsum, int M, int D,
                                   favit_stream stream) {
  FAVIT_CHECK_ARG(dy && x && mean && rstd && gamma && dx, "layernorm_bwd: null pointer");
  FAVIT_CHECK_ARG((dgamma == nullptr) == (dbeta == nullptr), "layernorm_bwd: dgamma and dbeta come together");
  FAVIT_CHECK_ARG(dxsum == nullptr || dgamma != nullptr, "layernorm_bwd: dxsum needs dgamma / dbeta");
  FAVIT_CHECK_ARG(M > 0 && D > 0, "layernorm_bwd: M, D must be positive");
  if (D % 4 != 0 || D > 128 * kMaxVec) {
    set_error("layernorm_bwd: D=%d unsupported (needs D %% 4 == 0 and D <= %d)", D, 128 * kMaxVec);
    return FAVIT_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int pairs = 2;                                          // rows in flight per CTA (2 warps each)
  const int nvh_t = D <= 256 ? 1 : (D <= 512 ? 2 : (D <= 768 ? 3 : 4));
  const size_t stage_bytes = (size_t)2 * 3 * nvh_t * 128 * 16;                       // <= 48 KiB
  const size_t smem = max(stage_bytes, (size_t)3 * pairs * D * sizeof(float));
// persistent grid = exactly the number of CTAs that are resident at once (queried, with the shared-memory carve-out
// maximised): a second, partial wave of a persistent kernel would run at a fraction of the occupancy
#define LN_BWD_NV(TX, TDY, NVH)                                                                                    \
  do {                                                                                                             \
    static int occ = 0;                                                                                            \
    if (occ == 0) {                                                                                                \
      FAVIT_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<TX, TDY, NVH>,                                           \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));              \
      FAVIT_CHECK_CUDA(cudaFuncSetAttribute(ln_bwd_kernel<TX, TDY, NVH>,                                           \
                                            cudaFuncAttributePreferredSharedMemoryCarveout,                        \
                                            cudaSharedmemCarveoutMaxShared));                                      \
      FAVIT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ln_bwd_kernel<TX, TDY, NVH>,            \
                                                                     pairs * 64, smem));                           \
      if (occ < 1) occ = 1;                                                                                        \
    }                                                                                                              \
    const unsigned blocks = (unsigned)max(1, min(ceil_div(M, pairs), occ * num_sms()));                            \
    ln_bwd_kernel<TX, TDY, NVH><<<blocks, pairs * 64, smem, st>>>((const TDY*)dy, (const TX*)x, mean, rstd, gamma, \
                                                                  dres, dx, (__nv_bfloat16*)dx_bf16, dgamma, dbeta, \
                                                                  dxsum, M, D);                                    \
  } while (0)
#define LN_BWD(TX, TDY)                           \
  do {                                            \
    if (D <= 256) LN_BWD_NV(TX, TDY, 1);          \
    else if (D <= 512) LN_BWD_NV(TX, TDY, 2);     \
    else if (D <= 768) LN_BWD_NV(TX, TDY, 3);     \
    else LN_BWD_NV(TX, TDY, 4);                   \
  } while (0)
  if (x_dtype == FAVIT_F32 && dy_dtype == FAVIT_BF16) LN_BWD(float, __nv_bfloat16);
  else if (x_dtype == FAVIT_F32 && dy_dtype == FAVIT_F32) LN_BWD(float, float);
  else if (x_dtype == FAVIT_BF16 && dy_dtype == FAVIT_BF16) LN_BWD(__nv_bfloat16, __nv_bfloat16);
  else if (x_dtype == FAVIT_BF16 && dy_dtype == FAVIT_F32) LN_BWD(__nv_bfloat16, float);
  else {
    set_error("layernorm_bwd: bad dtype");
    return FAVIT_ERR_ARG;
  }
#undef LN_BWD
#undef LN_BWD_NV
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}
