// Superpixel Patch Pooling for sm_100a: patch -> superpixel assignment and segment mean (fwd + bwd).
//
// Replaces /root/reference/models/sppp.py:91-128 (PatchToSuperpixelMapper.map_patches: g*g torch.unique calls
// with a device sync each) and :192-223 (SuperpixelPooling.pool 'mean': a Python loop of fancy-index + mean per
// superpixel), plus the per-image loop and torch.stack of models/sppp_mhla.py:286-300 — one launch per batch.
//
// Integer semantics reproduced bit for bit:
//   dominant label  = most frequent label of the patch, ties -> smallest label id   (sppp.py:117-120)
//   slot of a patch = rank of its dominant label by FIRST APPEARANCE in raster patch order, i.e. the
//                     insertion order of the reference's dict                        (sppp.py:123-126)
//   per-slot patch lists are ascending patch ids                                     (sppp.py:126)
//
// All three kernels are HBM-bound byte movers: labels are read once with full 128-byte lines per patch row,
// embeddings are read once with 16-byte vector loads through a CSR of patches per slot (no atomics, fixed
// summation order), gradients are a pure gather.
#include <limits.h>

#include "favit_common.cuh"

namespace favit {
namespace {

__device__ __forceinline__ long long warp_min_i64(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other < v ? other : v;
  }
  return v;
}

// One warp per patch.  PPL = pixels held per lane (ceil(ps*ps/32) rounded up to a power of two).
template <int PPL>
__global__ void __launch_bounds__(256) sppp_dominant_kernel(const int64_t* __restrict__ labels,
                                                            int64_t* __restrict__ dom, int B, int img_h,
                                                            int img_w, int ps, int grid) {
  const int P = grid * grid;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (int64_t)B * P) return;  // warp-uniform
  const int lane = threadIdx.x & 31;
  const int b = (int)(warp / P);
  const int p = (int)(warp % P);
  const int pi = p / grid, pj = p % grid;
  const int n = ps * ps;
  const int64_t* src = labels + ((int64_t)b * img_h + (int64_t)pi * ps) * img_w + (int64_t)pj * ps;

  long long vals[PPL];
  unsigned alive = 0;
#pragma unroll
  for (int t = 0; t < PPL; ++t) {
    const int idx = t * 32 + lane;
    vals[t] = LLONG_MAX;
    if (idx < n) {
      vals[t] = src[(int64_t)(idx / ps) * img_w + (idx % ps)];
      alive |= 1u << t;
    }
  }
  int remaining = n, best_cnt = 0;
  long long best_label = 0;
  while (remaining > best_cnt) {
    long long lmin = LLONG_MAX;
#pragma unroll
    for (int t = 0; t < PPL; ++t)
      if ((alive >> t) & 1u) lmin = vals[t] < lmin ? vals[t] : lmin;
    const long long cand = warp_min_i64(lmin);
    int c = 0;
#pragma unroll
    for (int t = 0; t < PPL; ++t)
      if (((alive >> t) & 1u) && vals[t] == cand) {
        ++c;
        alive &= ~(1u << t);
      }
    c = __reduce_add_sync(0xffffffffu, c);
    if (c > best_cnt) {
      best_cnt = c;
      best_label = cand;
    }
    remaining -= c;
  }
  if (lane == 0) dom[warp] = best_label;
}

// One CTA per image: first-seen slot ids, counts and the CSR (offsets/order) of patches per slot.
template <int T>
__global__ void __launch_bounds__(T) sppp_slot_kernel(const int64_t* __restrict__ dom, int32_t* __restrict__ slot,
                                                      int32_t* __restrict__ num_slots,
                                                      int32_t* __restrict__ counts,
                                                      int64_t* __restrict__ slot_label,
                                                      int32_t* __restrict__ offsets, int32_t* __restrict__ order,
                                                      int P, int r_cap) {
  constexpr int NW = T / 32;
  __shared__ int s_red[NW];
  __shared__ int s_scan[NW];
  __shared__ int s_bcast[2];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t* d = dom + (int64_t)b * P;
  int32_t* sl = slot + (int64_t)b * P;
  int32_t* ord = order + (int64_t)b * P;
  const int chunk = (P + T - 1) / T;
  const int p0 = min(P, tid * chunk), p1 = min(P, p0 + chunk);
  for (int p = p0; p < p1; ++p) sl[p] = -1;
  int first = p0;
  int r = 0, off = 0;
  while (true) {
    while (first < p1 && sl[first] >= 0) ++first;
    int cand = first < p1 ? first : INT_MAX;
    cand = __reduce_min_sync(0xffffffffu, cand);
    if (lane == 0) s_red[wid] = cand;
    __syncthreads();
    if (wid == 0) {
      int v = lane < NW ? s_red[lane] : INT_MAX;
      v = __reduce_min_sync(0xffffffffu, v);
      if (lane == 0) s_bcast[0] = v;
    }
    __syncthreads();
    const int pstar = s_bcast[0];
    if (pstar == INT_MAX) break;  // block-uniform
    const int64_t L = d[pstar];
    int cnt = 0;
    for (int p = max(p0, pstar); p < p1; ++p) cnt += (sl[p] < 0 && d[p] == L) ? 1 : 0;
    // block exclusive scan of cnt
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_scan[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = lane < NW ? s_scan[lane] : 0;
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += y;
      }
      if (lane < NW) s_scan[lane] = winc - w;  // exclusive warp offsets
      if (lane == 31) s_bcast[1] = winc;       // block total
    }
    __syncthreads();
    int pos = off + s_scan[wid] + incl - cnt;
    const int total = s_bcast[1];
    if (cnt > 0) {
      for (int p = max(p0, pstar); p < p1; ++p)
        if (sl[p] < 0 && d[p] == L) {
          sl[p] = r;
          ord[pos++] = p;
        }
    }
    if (tid == 0 && r < r_cap) {
      counts[(int64_t)b * r_cap + r] = total;
      slot_label[(int64_t)b * r_cap + r] = L;
      offsets[(int64_t)b * (r_cap + 1) + r] = off;
    }
    off += total;
    ++r;
    __syncthreads();  // s_bcast / s_red are rewritten next round
  }
  if (tid == 0) num_slots[b] = r;
  // tail: offsets of unused rows = P, counts = 0
  for (int q = r + tid; q <= r_cap; q += T) {
    if (q <= r_cap) offsets[(int64_t)b * (r_cap + 1) + q] = off;
    if (q < r_cap && q >= r) {
      counts[(int64_t)b * r_cap + q] = 0;
      slot_label[(int64_t)b * r_cap + q] = 0;
    }
  }
}

template <typename T> struct Vec8 { static constexpr bool ok = true; };

// One warp per (image, slot, 256-element column chunk).
template <typename TIn, typename TOut, int VEC>
__global__ void __launch_bounds__(256) sppp_pool_fwd_kernel(const TIn* __restrict__ x,
                                                            const int32_t* __restrict__ order,
                                                            const int32_t* __restrict__ offsets,
                                                            const int32_t* __restrict__ num_slots,
                                                            TOut* __restrict__ out, int B, int P, int R, int D,
                                                            int r_cap) {
  const int lane = threadIdx.x & 31;
  const int nchunks = (D + 32 * VEC - 1) / (32 * VEC);
  const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (item >= (int64_t)B * R * nchunks) return;
  const int dc = (int)(item % nchunks);
  const int r = (int)((item / nchunks) % R);
  const int b = (int)(item / ((int64_t)nchunks * R));
  const int d0 = dc * 32 * VEC + lane * VEC;
  if (d0 >= D) return;
  float acc[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
  int beg = 0, end = 0;
  if (r < r_cap && r < num_slots[b]) {
    beg = offsets[(int64_t)b * (r_cap + 1) + r];
    end = offsets[(int64_t)b * (r_cap + 1) + r + 1];
  }
  const int32_t* ord = order + (int64_t)b * P;
  const TIn* xb = x + (int64_t)b * P * D + d0;
  int t = beg;
  if constexpr (VEC == 8) {
    for (; t + 4 <= end; t += 4) {
      float f[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8(xb + (int64_t)ord[t + u] * D, f[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[u][e];
    }
    for (; t < end; ++t) {
      float f[8];
      load8(xb + (int64_t)ord[t] * D, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  } else {
    for (; t < end; ++t) acc[0] += Elem<TIn>::ld(xb + (int64_t)ord[t] * D);
  }
  const float n = (float)max(end - beg, 1);
  TOut* o = out + ((int64_t)b * R + r) * D + d0;
  if constexpr (VEC == 8) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = acc[e] / n;
    store8(o, f);
  } else {
    Elem<TOut>::st(o, acc[0] / n);
  }
}

// One warp per (image, patch, column chunk): dx[b,p,:] = dout[b,slot[b,p],:] / counts[b,slot[b,p]].
template <typename TIn, typename TOut, int VEC>
__global__ void __launch_bounds__(256) sppp_pool_bwd_kernel(const TIn* __restrict__ dout,
                                                            const int32_t* __restrict__ slot,
                                                            const int32_t* __restrict__ counts,
                                                            TOut* __restrict__ dx, int B, int P, int R, int D,
                                                            int r_cap) {
  const int lane = threadIdx.x & 31;
  const int nchunks = (D + 32 * VEC - 1) / (32 * VEC);
  const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (item >= (int64_t)B * P * nchunks) return;
  const int dc = (int)(item % nchunks);
  const int p = (int)((item / nchunks) % P);
  const int b = (int)(item / ((int64_t)nchunks * P));
  const int d0 = dc * 32 * VEC + lane * VEC;
  if (d0 >= D) return;
  const int r = slot[(int64_t)b * P + p];
  TOut* o = dx + ((int64_t)b * P + p) * D + d0;
  const bool live = r >= 0 && r < R && r < r_cap;
  const float n = live ? (float)max(counts[(int64_t)b * r_cap + r], 1) : 1.f;
  if constexpr (VEC == 8) {
    float f[8];
    if (live) {
      load8(dout + ((int64_t)b * R + r) * D + d0, f);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = f[e] / n;
    store8(o, f);
  } else {
    const float v = live ? Elem<TIn>::ld(dout + ((int64_t)b * R + r) * D + d0) : 0.f;
    Elem<TOut>::st(o, v / n);
  }
}

template <typename TIn, typename TOut>
int launch_pool_fwd(const void* x, const int32_t* order, const int32_t* offsets, const int32_t* num_slots,
                    void* out, int B, int P, int R, int D, int r_cap, cudaStream_t st) {
  const bool vec = (D % 8 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
  const int per = vec ? 256 : 32;
  const int64_t items = (int64_t)B * R * ceil_div(D, per);
  const unsigned blocks = (unsigned)ceil_div64(items, 8);
  if (vec)
    sppp_pool_fwd_kernel<TIn, TOut, 8><<<blocks, 256, 0, st>>>((const TIn*)x, order, offsets, num_slots,
                                                                (TOut*)out, B, P, R, D, r_cap);
  else
    sppp_pool_fwd_kernel<TIn, TOut, 1><<<blocks, 256, 0, st>>>((const TIn*)x, order, offsets, num_slots,
                                                                (TOut*)out, B, P, R, D, r_cap);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

template <typename TIn, typename TOut>
int launch_pool_bwd(const void* dout, const int32_t* slot, const int32_t* counts, void* dx, int B, int P, int R,
                    int D, int r_cap, cudaStream_t st) {
  const bool vec = (D % 8 == 0) && ((uintptr_t)dout % 16 == 0) && ((uintptr_t)dx % 16 == 0);
  const int per = vec ? 256 : 32;
  const int64_t items = (int64_t)B * P * ceil_div(D, per);
  const unsigned blocks = (unsigned)ceil_div64(items, 8);
  if (vec)
    sppp_pool_bwd_kernel<TIn, TOut, 8><<<blocks, 256, 0, st>>>((const TIn*)dout, slot, counts, (TOut*)dx, B, P,
                                                                R, D, r_cap);
  else
    sppp_pool_bwd_kernel<TIn, TOut, 1><<<blocks, 256, 0, st>>>((const TIn*)dout, slot, counts, (TOut*)dx, B, P,
                                                                R, D, r_cap);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

}  // namespace
}  // namespace favit

using namespace favit;

extern "C" int favit_sppp_assign(const int64_t* labels, int B, int img_h, int img_w, int patch, int grid,
                                 int64_t* dom, int32_t* slot, int32_t* num_slots, int32_t* counts,
                                 int64_t* slot_label, int32_t* offsets, int32_t* order, int r_cap,
                                 favit_stream stream) {
  FAVIT_CHECK_ARG(labels && dom && slot && num_slots && counts && slot_label && offsets && order,
                  "sppp_assign: null pointer");
  FAVIT_CHECK_ARG(B > 0 && patch > 0 && grid > 0 && r_cap > 0, "sppp_assign: B, patch, grid, r_cap must be > 0");
  FAVIT_CHECK_ARG((int64_t)grid * patch <= img_h && (int64_t)grid * patch <= img_w,
                  "sppp_assign: grid*patch (%d) exceeds the label map (%dx%d)", grid * patch, img_h, img_w);
  if (patch > 32) {
    set_error("sppp_assign: patch_size %d > 32 unsupported", patch);
    return FAVIT_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int P = grid * grid;
  const int ppl = ceil_div(patch * patch, 32);
  const unsigned blocks = (unsigned)ceil_div64((int64_t)B * P, 8);
#define FAVIT_DOM(PPL)                                                                                   \
  sppp_dominant_kernel<PPL><<<blocks, 256, 0, st>>>(labels, dom, B, img_h, img_w, patch, grid)
  if (ppl <= 1) FAVIT_DOM(1);
  else if (ppl <= 2) FAVIT_DOM(2);
  else if (ppl <= 4) FAVIT_DOM(4);
  else if (ppl <= 8) FAVIT_DOM(8);
  else if (ppl <= 16) FAVIT_DOM(16);
  else FAVIT_DOM(32);
#undef FAVIT_DOM
  FAVIT_CHECK_LAUNCH();
  if (P <= 1024)
    sppp_slot_kernel<256><<<B, 256, 0, st>>>(dom, slot, num_slots, counts, slot_label, offsets, order, P, r_cap);
  else
    sppp_slot_kernel<1024><<<B, 1024, 0, st>>>(dom, slot, num_slots, counts, slot_label, offsets, order, P, r_cap);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

extern "C" int favit_sppp_pool_fwd(const void* x, favit_dtype x_dtype, const int32_t* order,
                                   const int32_t* offsets, const int32_t* num_slots, void* out,
                                   favit_dtype out_dtype, int B, int P, int R, int D, int r_cap,
                                   favit_stream stream) {
  FAVIT_CHECK_ARG(x && order && offsets && num_slots && out, "sppp_pool_fwd: null pointer");
  FAVIT_CHECK_ARG(B > 0 && P > 0 && R > 0 && D > 0 && r_cap > 0, "sppp_pool_fwd: sizes must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == FAVIT_BF16 && out_dtype == FAVIT_F32)
    return launch_pool_fwd<__nv_bfloat16, float>(x, order, offsets, num_slots, out, B, P, R, D, r_cap, st);
  if (x_dtype == FAVIT_BF16 && out_dtype == FAVIT_BF16)
    return launch_pool_fwd<__nv_bfloat16, __nv_bfloat16>(x, order, offsets, num_slots, out, B, P, R, D, r_cap, st);
  if (x_dtype == FAVIT_F32 && out_dtype == FAVIT_F32)
    return launch_pool_fwd<float, float>(x, order, offsets, num_slots, out, B, P, R, D, r_cap, st);
  if (x_dtype == FAVIT_F32 && out_dtype == FAVIT_BF16)
    return launch_pool_fwd<float, __nv_bfloat16>(x, order, offsets, num_slots, out, B, P, R, D, r_cap, st);
  set_error("sppp_pool_fwd: bad dtype");
  return FAVIT_ERR_ARG;
}

extern "C" int favit_sppp_pool_bwd(const void* dout, favit_dtype dout_dtype, const int32_t* slot,
                                   const int32_t* counts, void* dx, favit_dtype dx_dtype, int B, int P, int R,
                                   int D, int r_cap, favit_stream stream) {
  FAVIT_CHECK_ARG(dout && slot && counts && dx, "sppp_pool_bwd: null pointer");
  FAVIT_CHECK_ARG(B > 0 && P > 0 && R > 0 && D > 0 && r_cap > 0, "sppp_pool_bwd: sizes must be > 0");
  cudaStream_t st = (cudaStream_t)stream;
  if (dout_dtype == FAVIT_F32 && dx_dtype == FAVIT_F32)
    return launch_pool_bwd<float, float>(dout, slot, counts, dx, B, P, R, D, r_cap, st);
  if (dout_dtype == FAVIT_F32 && dx_dtype == FAVIT_BF16)
    return launch_pool_bwd<float, __nv_bfloat16>(dout, slot, counts, dx, B, P, R, D, r_cap, st);
  if (dout_dtype == FAVIT_BF16 && dx_dtype == FAVIT_BF16)
    return launch_pool_bwd<__nv_bfloat16, __nv_bfloat16>(dout, slot, counts, dx, B, P, R, D, r_cap, st);
  if (dout_dtype == FAVIT_BF16 && dx_dtype == FAVIT_F32)
    return launch_pool_bwd<__nv_bfloat16, float>(dout, slot, counts, dx, B, P, R, D, r_cap, st);
  set_error("sppp_pool_bwd: bad dtype");
  return FAVIT_ERR_ARG;
}
