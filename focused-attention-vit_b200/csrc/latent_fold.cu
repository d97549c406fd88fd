// Latent-projection fold for MHLA (sm_100a), forward and backward.
//
// /root/reference/models/mhla.py:105-106 applies one shared Linear(hd, hd) ("latent_proj": Wl, bl) to every key and
// value vector.  Algebraically (verified against the reference to 1e-15 in fp64) that is
//     K path:  q.(Wl k + bl) = (Wl^T q).k + q.bl, and q.bl is constant along the softmax axis      -> fold into Wq, bq
//     V path:  sum_j p_j (Wl v_j + bl) = Wl (sum_j p_j v_j) + bl   because softmax rows sum to one -> fold into Wp, bp
// so the hot path never touches K/V a second time:
//     Wq'[h,a,:] = sum_b Wl[b,a] Wq[h,b,:]     bq'[h,a] = sum_b bq[h,b] Wl[b,a]
//     Wp'[o,h,a] = sum_b Wp[o,h,b] Wl[b,a]     bp'[o]   = bp[o] + sum_{h,b} Wp[o,h,b] bl[b]
// These kernels produce the folded weights directly in the GEMM operand dtype (one pass, no torch.cat / matmul /
// cast chain) and, in backward, map the gradients of the folded weights back to qkv / proj / latent_proj:
//     dWq[h,b,:] = sum_a Wl[b,a] dWq'[h,a,:]               dbq[h,b] = sum_a dbq'[h,a] Wl[b,a]
//     dWp[o,h,b] = sum_a dWp'[o,h,a] Wl[b,a] + dbp'[o] bl[b]        dbp = dbp'
//     dWl[b,a]   = sum_{h,c} Wq[h,b,c] dWq'[h,a,c] + sum_h bq[h,b] dbq'[h,a] + sum_{o,h} Wp[o,h,b] dWp'[o,h,a]
//     dbl[b]     = sum_{o,h} dbp'[o] Wp[o,h,b]
// All fp32 FFMA (O(D^2 hd) work, independent of the number of tokens); head_dim <= 64.
#include "favit_common.cuh"

namespace favit {
namespace {

constexpr int kTile = 64;  // columns (q fold) or rows (proj fold) per CTA

template <typename T> __device__ __forceinline__ void put(T* p, float v);
template <> __device__ __forceinline__ void put<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void put<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }


__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// Wl (hd x hd, hd % 4 == 0, hd <= 64) -> shared memory, optionally transposed.  All global loads of a thread are issued
// before the first use (these kernels are tiny and otherwise pay one DRAM latency per load).
template <bool TRANS>
__device__ __forceinline__ void load_lat(float* sL, const float* __restrict__ lat_w, int hd) {
  float4 v[4];
  const int n4 = hd * hd / 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = threadIdx.x + 256 * k;
    if (idx < n4) v[k] = ld4(lat_w + 4 * idx);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = threadIdx.x + 256 * k;
    if (idx < n4) {
      if (!TRANS) {
        *reinterpret_cast<float4*>(sL + 4 * idx) = v[k];
      } else {
        const int r = (4 * idx) / hd, c = (4 * idx) % hd;  // element (r, c..c+3) of Wl goes to sL[c..][r]
        sL[c * hd + r] = v[k].x; sL[(c + 1) * hd + r] = v[k].y; sL[(c + 2) * hd + r] = v[k].z; sL[(c + 3) * hd + r] = v[k].w;
      }
    }
  }
}

// out[h,a,c] = sum_b M(b,a) in[h,b,c] for one head h and a 64-column chunk; M = Wl (TRANS = false) or Wl^T.
// `in` and `out` may be the same buffer (each CTA reads its whole tile before writing it).
// The first column chunk of every head also folds that head's bias: bias_out[h,a] = sum_b bias_in[h,b] M(b,a)
// (bias_in / bias_out may alias).
template <typename TOut, bool TRANS>
__device__ __forceinline__ void fold_q_body(const float* in, TOut* out, const float* __restrict__ lat_w,
                                            const float* bias_in, float* bias_out, int H, int hd, int D) {
  extern __shared__ float sm[];
  __shared__ float s_vec[64];
  float* sL = sm;             // [hd][hd]
  float* sT = sm + hd * hd;   // [hd][kTile]
  const int h = blockIdx.y, c0 = blockIdx.x * kTile;
  // sL[b][a] = M(b,a): transposed while loading so that the inner loop reads consecutive / broadcast words
  {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = threadIdx.x + 256 * k;   // float4 index in the [hd][64] tile
      const int b = idx / (kTile / 4), c = (idx % (kTile / 4)) * 4;
      v[k] = (idx < hd * kTile / 4 && c0 + c < D) ? ld4(in + ((int64_t)h * hd + b) * D + c0 + c) : make_float4(0, 0, 0, 0);
    }
    load_lat<TRANS>(sL, lat_w, hd);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = threadIdx.x + 256 * k;
      if (idx < hd * kTile / 4) *reinterpret_cast<float4*>(sT + 4 * idx) = v[k];
    }
    if (bias_in && blockIdx.x == 0 && threadIdx.x < hd) s_vec[threadIdx.x] = bias_in[h * hd + threadIdx.x];
  }
  __syncthreads();
  if (bias_in && blockIdx.x == 0 && threadIdx.x < hd) {
    float acc = 0.f;
    for (int b = 0; b < hd; ++b) acc = fmaf(s_vec[b], sL[b * hd + threadIdx.x], acc);
    bias_out[h * hd + threadIdx.x] = acc;
  }
  // 4 x 4 outputs per thread: 16 independent FMA chains per pair of 16-byte shared loads
  const int a4 = (threadIdx.x / (kTile / 4)) * 4, c4 = (threadIdx.x % (kTile / 4)) * 4;
  if (a4 < hd) {
    float acc[4][4] = {};
#pragma unroll 4
    for (int b = 0; b < hd; ++b) {
      const float4 l = *reinterpret_cast<const float4*>(sL + b * hd + a4);
      const float4 t = *reinterpret_cast<const float4*>(sT + b * kTile + c4);
      const float lv[4] = {l.x, l.y, l.z, l.w}, tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(lv[i], tv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + c4 + j < D) put(out + ((int64_t)h * hd + a4 + i) * D + c0 + c4 + j, acc[i][j]);
  }
}

// out[o,h,a] = sum_b in[o,h,b] M(b,a) (+ rowscale[o] * colvec[a]) for 64 rows o and one head h.
// With bias_out != nullptr (forward) it also adds this head's share of sum_b in[o,h,b] bias_vec[b] to bias_out[o].
template <typename TOut, bool TRANS>
__device__ __forceinline__ void fold_p_body(const float* in, TOut* out, const float* __restrict__ lat_w,
                                            const float* __restrict__ rowscale, const float* __restrict__ colvec,
                                            float* __restrict__ bias_out, const float* __restrict__ bias_vec, int H,
                                            int hd, int D) {
  extern __shared__ float sm[];
  float* sL = sm;             // [hd][hd]
  float* sT = sm + hd * hd;   // [kTile][hd]
  const int h = blockIdx.y, o0 = blockIdx.x * kTile;
  {
    float4 v[4];
    const int g = hd / 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = threadIdx.x + 256 * k;   // float4 index in the [64][hd] tile
      const int o = idx / g, b = (idx % g) * 4;
      v[k] = (idx < kTile * g && o0 + o < D) ? ld4(in + (int64_t)(o0 + o) * D + h * hd + b) : make_float4(0, 0, 0, 0);
    }
    load_lat<TRANS>(sL, lat_w, hd);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = threadIdx.x + 256 * k;
      if (idx < kTile * g) *reinterpret_cast<float4*>(sT + 4 * idx) = v[k];
    }
  }
  __syncthreads();
  {
    const int ag = hd / 4;
    const int o4 = (threadIdx.x / ag) * 4, a4 = (threadIdx.x % ag) * 4;
    if (o4 < kTile) {
      float acc[4][4] = {};
#pragma unroll 4
      for (int b = 0; b < hd; ++b) {
        const float4 l = *reinterpret_cast<const float4*>(sL + b * hd + a4);
        const float lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float tv = sT[(o4 + i) * hd + b];  // broadcast within the warp
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(tv, lv[j], acc[i][j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int o = o0 + o4 + i;
        if (o < D) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v = acc[i][j];
            if (rowscale) v = fmaf(rowscale[o], colvec[a4 + j], v);
            put(out + (int64_t)o * D + h * hd + a4 + j, v);
          }
        }
      }
    }
  }
  if (bias_out) {
    for (int o = threadIdx.x; o < kTile; o += blockDim.x) {
      if (o0 + o >= D) continue;
      float acc = 0.f;
      for (int b = 0; b < hd; ++b) acc = fmaf(sT[o * hd + b], bias_vec[b], acc);
      atomicAdd(bias_out + o0 + o, acc);
    }
  }
}

// forward odds and ends: K/V weight rows cast, folded q bias, K/V bias copy, folded proj bias.
template <typename TOut>
__device__ __forceinline__ void fold_misc_fwd_body(const float* __restrict__ qkv_w, const float* __restrict__ qkv_b,
                                                   const float* __restrict__ proj_b, TOut* __restrict__ wqkv_c,
                                                   float* __restrict__ bqkv, float* __restrict__ bproj, int D) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nkv = (int64_t)2 * D * D;
  for (int64_t i = tid * 4; i < nkv; i += stride * 4) {   // D*D is a multiple of 4 (hd % 4 == 0)
    const float4 v = *reinterpret_cast<const float4*>(qkv_w + (int64_t)D * D + i);
    TOut* o = wqkv_c + (int64_t)D * D + i;
    put(o, v.x); put(o + 1, v.y); put(o + 2, v.z); put(o + 3, v.w);
  }
  for (int64_t i = D + tid; i < 3 * D; i += stride) bqkv[i] = qkv_b[i];   // the q part comes from fold_q_kernel
  for (int64_t o = tid; o < D; o += stride) bproj[o] = proj_b[o];  // fold_p_kernel adds sum_{h,b} Wp[o,h,b] bl[b]
}

// dWl partial sums.  blockIdx.z == 0: sum over kChunks 64-column chunks of head h of Wq[h,b,c] dWq'[h,a,c] (+ the
// bias term once per head); blockIdx.z == 1: sum over 64-row chunks of Wp[o,h,b] dWp'[o,h,a], and dbl.
constexpr int kChunks = 4;
__device__ __forceinline__ void fold_dlat_body(const float* __restrict__ qkv_w, const float* __restrict__ qkv_b,
                                               const float* __restrict__ proj_w, const float* __restrict__ dwq,
                                               const float* __restrict__ dbq, const float* __restrict__ dwp,
                                               const float* __restrict__ dbp, float* __restrict__ dlat_w,
                                               float* __restrict__ dlat_b, int H, int hd, int D, bool proj) {
  extern __shared__ float sm[];
  float* sA = sm;                // q: [kTile c][hd b]   proj: [kTile o][hd b]   (weight tile)
  float* sB = sm + hd * kTile;   // q: [kTile c][hd a]   proj: [kTile o][hd a]   (gradient tile)
  const int h = blockIdx.y;
  const int ag = hd / 4;
  const int b4 = (threadIdx.x / ag) * 4, a4 = (threadIdx.x % ag) * 4;   // this thread's 4 x 4 block of dWl
  const bool live = b4 < hd;
  float acc[4][4] = {};
  float accb = 0.f;
  for (int ch = 0; ch < kChunks; ++ch) {
    const int t0 = (blockIdx.x * kChunks + ch) * kTile;
    if (t0 >= D) break;
    __syncthreads();
    {
      float4 va[4], vb[4];
      const int n4 = hd * kTile / 4, g = hd / 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = threadIdx.x + 256 * k;
        va[k] = vb[k] = make_float4(0, 0, 0, 0);
        if (idx < n4) {
          if (!proj) {
            const int b = idx / (kTile / 4), c = (idx % (kTile / 4)) * 4;   // coalesced along c
            if (t0 + c < D) {
              va[k] = ld4(qkv_w + ((int64_t)h * hd + b) * D + t0 + c);
              vb[k] = ld4(dwq + ((int64_t)h * hd + b) * D + t0 + c);
            }
          } else {
            const int o = idx / g, b = (idx % g) * 4;
            if (t0 + o < D) {
              va[k] = ld4(proj_w + (int64_t)(t0 + o) * D + h * hd + b);
              vb[k] = ld4(dwp + (int64_t)(t0 + o) * D + h * hd + b);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = threadIdx.x + 256 * k;
        if (idx < n4) {
          if (!proj) {  // stored transposed: [c][b]
            const int b = idx / (kTile / 4), c = (idx % (kTile / 4)) * 4;
            sA[c * hd + b] = va[k].x; sA[(c + 1) * hd + b] = va[k].y; sA[(c + 2) * hd + b] = va[k].z; sA[(c + 3) * hd + b] = va[k].w;
            sB[c * hd + b] = vb[k].x; sB[(c + 1) * hd + b] = vb[k].y; sB[(c + 2) * hd + b] = vb[k].z; sB[(c + 3) * hd + b] = vb[k].w;
          } else {
            *reinterpret_cast<float4*>(sA + 4 * idx) = va[k];
            *reinterpret_cast<float4*>(sB + 4 * idx) = vb[k];
          }
        }
      }
    }
    __syncthreads();
    if (live) {
#pragma unroll 4
      for (int t = 0; t < kTile; ++t) {
        const float4 x = *reinterpret_cast<const float4*>(sA + t * hd + b4);
        const float4 y = *reinterpret_cast<const float4*>(sB + t * hd + a4);
        const float xv[4] = {x.x, x.y, x.z, x.w}, yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
      }
    }
    if (proj && threadIdx.x < hd) {
      const int b = threadIdx.x;
      for (int o = 0; o < kTile; ++o)
        if (t0 + o < D) accb = fmaf(dbp[t0 + o], sA[o * hd + b], accb);
    }
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[i][j];
        if (!proj && blockIdx.x == 0) v = fmaf(qkv_b[h * hd + b4 + i], dbq[h * hd + a4 + j], v);
        atomicAdd(dlat_w + (b4 + i) * hd + a4 + j, v);
      }
  }
  if (proj && threadIdx.x < hd) atomicAdd(dlat_b + threadIdx.x, accb);
}

// dbq[h,b] = sum_a dbq'[h,a] Wl[b,a], in place on the first D entries of the folded bias gradient.
__global__ void __launch_bounds__(256) fold_dbq_kernel(float* dbqkv, const float* __restrict__ lat_w, int H, int hd) {
  extern __shared__ float sm[];  // [H*hd] copy of dbq'
  const int D = H * hd;
  for (int i = threadIdx.x; i < D; i += blockDim.x) sm[i] = dbqkv[i];
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const int h = i / hd, b = i % hd;
    float acc = 0.f;
    for (int a = 0; a < hd; ++a) acc = fmaf(sm[h * hd + a], lat_w[b * hd + a], acc);
    dbqkv[i] = acc;
  }
}


// ---- batched over layers ----------------------------------------------------------------------------------------
// The fold is O(D^2 hd) weight-only work; a launch per layer costs more in launch latency than in arithmetic, so every
// kernel takes the pointers of up to kMaxLayers layers by value (blockIdx.z selects the layer) and one launch serves
// the whole model.  Pointer order per layer: see favit_latent_fold_fwd_batched / _bwd_batched in include/favit.h.
constexpr int kMaxLayers = 16;
constexpr int kFwdPtrs = 10, kBwdPtrs = 11;
struct FwdTable { const void* p[kMaxLayers][kFwdPtrs]; };
struct BwdTable { const void* p[kMaxLayers][kBwdPtrs]; };
#define FP(k) reinterpret_cast<const float*>(t.p[blockIdx.z][k])
#define FM(k) const_cast<float*>(reinterpret_cast<const float*>(t.p[blockIdx.z][k]))

template <typename TOut>
__global__ void __launch_bounds__(256) fold_misc_fwd_kernel(const __grid_constant__ FwdTable t, int D) {
  fold_misc_fwd_body<TOut>(FP(0), FP(1), FP(3), reinterpret_cast<TOut*>(const_cast<void*>(t.p[blockIdx.z][6])), FM(7),
                           FM(9), D);
}
template <typename TOut>
__global__ void __launch_bounds__(256) fold_q_fwd_kernel(const __grid_constant__ FwdTable t, int H, int hd, int D) {
  fold_q_body<TOut, false>(FP(0), reinterpret_cast<TOut*>(const_cast<void*>(t.p[blockIdx.z][6])), FP(4), FP(1), FM(7), H,
                           hd, D);
}
template <typename TOut>
__global__ void __launch_bounds__(256) fold_p_fwd_kernel(const __grid_constant__ FwdTable t, int H, int hd, int D) {
  fold_p_body<TOut, false>(FP(2), reinterpret_cast<TOut*>(const_cast<void*>(t.p[blockIdx.z][8])), FP(4), nullptr, nullptr,
                           FM(9), FP(5), H, hd, D);
}
#undef FP
#undef FM
#define BP(k) reinterpret_cast<const float*>(t.p[layer][k])
#define BM(k) const_cast<float*>(reinterpret_cast<const float*>(t.p[layer][k]))
__global__ void __launch_bounds__(256) fold_dlat_kernel(const __grid_constant__ BwdTable t, int H, int hd, int D) {
  const int layer = blockIdx.z >> 1;
  fold_dlat_body(BP(0), BP(1), BP(2), BP(5), BP(6), BP(7), BP(8), BM(9), BM(10), H, hd, D, (blockIdx.z & 1) != 0);
}
__global__ void __launch_bounds__(256) fold_q_bwd_kernel(const __grid_constant__ BwdTable t, int H, int hd, int D) {
  const int layer = blockIdx.z;
  fold_q_body<float, true>(BP(5), BM(5), BP(3), BP(6), BM(6), H, hd, D);
}
__global__ void __launch_bounds__(256) fold_p_bwd_kernel(const __grid_constant__ BwdTable t, int H, int hd, int D) {
  const int layer = blockIdx.z;
  fold_p_body<float, true>(BP(7), BM(7), BP(3), BP(8), BP(4), nullptr, nullptr, H, hd, D);
}
#undef BP
#undef BM

int check(int H, int hd, const char* who) {
  FAVIT_CHECK_ARG(H > 0 && hd > 0, "%s: H, hd must be positive", who);
  if (hd > 64 || hd % 4 != 0) {
    set_error("%s: head_dim %d unsupported by the fused fold (needs hd <= 64, hd %% 4 == 0)", who, hd);
    return FAVIT_ERR_UNSUPPORTED;
  }
  return FAVIT_OK;
}

}  // namespace
}  // namespace favit

using namespace favit;

extern "C" int favit_latent_fold_fwd_batched(int L, const void* const* ptrs, int H, int hd, favit_dtype out_dtype,
                                             favit_stream stream) {
  FAVIT_CHECK_ARG(L > 0 && ptrs, "latent_fold_fwd_batched: L must be > 0 and ptrs non-null");
  if (int rc = check(H, hd, "latent_fold_fwd")) return rc;
  FAVIT_CHECK_ARG(out_dtype == FAVIT_BF16 || out_dtype == FAVIT_F32, "latent_fold_fwd: bad dtype");
  for (int i = 0; i < L * kFwdPtrs; ++i) FAVIT_CHECK_ARG(ptrs[i], "latent_fold_fwd: null pointer (entry %d)", i);
  cudaStream_t st = (cudaStream_t)stream;
  const int D = H * hd;
  const size_t smem = (size_t)(hd * hd + hd * kTile) * sizeof(float);
  for (int l0 = 0; l0 < L; l0 += kMaxLayers) {
    const int n = L - l0 < kMaxLayers ? L - l0 : kMaxLayers;
    FwdTable t;
    for (int l = 0; l < n; ++l)
      for (int k = 0; k < kFwdPtrs; ++k) t.p[l][k] = ptrs[(size_t)(l0 + l) * kFwdPtrs + k];
    const dim3 gm(2 * num_sms() / n + 1, 1, n), gq(ceil_div(D, kTile), H, n);
    if (out_dtype == FAVIT_BF16) {
      fold_misc_fwd_kernel<__nv_bfloat16><<<gm, 256, 0, st>>>(t, D);
      FAVIT_CHECK_LAUNCH();
      fold_q_fwd_kernel<__nv_bfloat16><<<gq, 256, smem, st>>>(t, H, hd, D);
      FAVIT_CHECK_LAUNCH();
      fold_p_fwd_kernel<__nv_bfloat16><<<gq, 256, smem, st>>>(t, H, hd, D);
      FAVIT_CHECK_LAUNCH();
    } else {
      fold_misc_fwd_kernel<float><<<gm, 256, 0, st>>>(t, D);
      FAVIT_CHECK_LAUNCH();
      fold_q_fwd_kernel<float><<<gq, 256, smem, st>>>(t, H, hd, D);
      FAVIT_CHECK_LAUNCH();
      fold_p_fwd_kernel<float><<<gq, 256, smem, st>>>(t, H, hd, D);
      FAVIT_CHECK_LAUNCH();
    }
  }
  return FAVIT_OK;
}

extern "C" int favit_latent_fold_fwd(const float* qkv_w, const float* qkv_b, const float* proj_w, const float* proj_b,
                                     const float* lat_w, const float* lat_b, void* wqkv_c, float* bqkv, void* wproj_c,
                                     float* bproj, int H, int hd, favit_dtype out_dtype, favit_stream stream) {
  const void* p[kFwdPtrs] = {qkv_w, qkv_b, proj_w, proj_b, lat_w, lat_b, wqkv_c, bqkv, wproj_c, bproj};
  return favit_latent_fold_fwd_batched(1, p, H, hd, out_dtype, stream);
}

// Per layer: dwqkv [3D,D], dbqkv [3D], dwproj [D,D] hold the gradients of the FOLDED weights on entry and the gradients
// of qkv.weight / qkv.bias / proj.weight on exit (transformed in place; the K/V rows and dbproj need no change).
// dlat_w [hd,hd] and dlat_b [hd] are overwritten.
extern "C" int favit_latent_fold_bwd_batched(int L, const void* const* ptrs, int H, int hd, favit_stream stream) {
  FAVIT_CHECK_ARG(L > 0 && ptrs, "latent_fold_bwd_batched: L must be > 0 and ptrs non-null");
  if (int rc = check(H, hd, "latent_fold_bwd")) return rc;
  for (int i = 0; i < L * kBwdPtrs; ++i) FAVIT_CHECK_ARG(ptrs[i], "latent_fold_bwd: null pointer (entry %d)", i);
  cudaStream_t st = (cudaStream_t)stream;
  const int D = H * hd;
  const size_t smem = (size_t)(hd * hd + hd * kTile) * sizeof(float);
  for (int l = 0; l < L; ++l) {
    float* dlw = const_cast<float*>(reinterpret_cast<const float*>(ptrs[(size_t)l * kBwdPtrs + 9]));
    float* dlb = const_cast<float*>(reinterpret_cast<const float*>(ptrs[(size_t)l * kBwdPtrs + 10]));
    if (dlb == dlw + hd * hd) {  // the host side keeps [dlat_w | dlat_b] of a layer (and of all layers) contiguous
      int run = 1;
      while (l + run < L && ptrs[(size_t)(l + run) * kBwdPtrs + 9] == (const void*)(dlw + (size_t)run * (hd * hd + hd)) &&
             ptrs[(size_t)(l + run) * kBwdPtrs + 10] == (const void*)(dlw + (size_t)run * (hd * hd + hd) + hd * hd))
        ++run;
      FAVIT_CHECK_CUDA(cudaMemsetAsync(dlw, 0, (size_t)run * (hd * hd + hd) * sizeof(float), st));
      l += run - 1;
    } else {
      FAVIT_CHECK_CUDA(cudaMemsetAsync(dlw, 0, (size_t)hd * hd * sizeof(float), st));
      FAVIT_CHECK_CUDA(cudaMemsetAsync(dlb, 0, (size_t)hd * sizeof(float), st));
    }
  }
  for (int l0 = 0; l0 < L; l0 += kMaxLayers) {
    const int n = L - l0 < kMaxLayers ? L - l0 : kMaxLayers;
    BwdTable t;
    for (int l = 0; l < n; ++l)
      for (int k = 0; k < kBwdPtrs; ++k) t.p[l][k] = ptrs[(size_t)(l0 + l) * kBwdPtrs + k];
    // latent gradients first: they need the gradients of the folded weights as they came in
    fold_dlat_kernel<<<dim3(ceil_div(ceil_div(D, kTile), kChunks), H, 2 * n), 256,
                       (size_t)2 * hd * kTile * sizeof(float), st>>>(t, H, hd, D);
    FAVIT_CHECK_LAUNCH();
    fold_q_bwd_kernel<<<dim3(ceil_div(D, kTile), H, n), 256, smem, st>>>(t, H, hd, D);
    FAVIT_CHECK_LAUNCH();
    fold_p_bwd_kernel<<<dim3(ceil_div(D, kTile), H, n), 256, smem, st>>>(t, H, hd, D);
    FAVIT_CHECK_LAUNCH();
  }
  return FAVIT_OK;
}

extern "C" int favit_latent_fold_bwd(const float* qkv_w, const float* qkv_b, const float* proj_w, const float* lat_w,
                                     const float* lat_b, float* dwqkv, float* dbqkv, float* dwproj, const float* dbproj,
                                     float* dlat_w, float* dlat_b, int H, int hd, favit_stream stream) {
  const void* p[kBwdPtrs] = {qkv_w, qkv_b, proj_w, lat_w, lat_b, dwqkv, dbqkv, dwproj, dbproj, dlat_w, dlat_b};
  return favit_latent_fold_bwd_batched(1, p, H, hd, stream);
}
