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
// All kernels are HBM-bound byte movers and are built around bytes in flight, not arithmetic:
//   dominant : one warp per patch, 16-byte label loads, one vote for single-label patches, else a count per
//              distinct label (few per patch)
//   slots    : one CTA per image in shared memory: first occurrence of each dominant label, prefix count of the
//              leaders, rank inside the slot -> slot ids, counts and the CSR, a handful of barriers per image
//   pool fwd : one CTA per (image, 128-byte column slice); the [P x 128 B] tile arrives by TMA
//              (cp.async.bulk.tensor, mbarrier) while the CSR is fetched, then each (slot, 8-byte column group)
//              thread sums its patches out of shared memory in ascending patch order (deterministic, no atomics)
//   pool bwd : one CTA per (image, 64-column slice); the R pre-divided gradient rows sit in shared memory and every
//              patch row is one 16-byte store per thread
// The older warp-per-item kernels remain as the fallback for shapes the tiled ones do not take (odd strides, huge R).
#include <cuda.h>
#include <limits.h>
#include <math_constants.h>

#include <algorithm>
#include <mutex>

#include "favit_common.cuh"
#include "tcgen05_ptx.cuh"

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

// =====================================================================================================
// Tiled kernels (the fast path)
// =====================================================================================================

// ---- dominant label, vector loads --------------------------------------------------------------------
// One warp per patch; PS in {8, 16, 32}: the PS*PS labels are held as NV = PS*PS/32 int64 per lane, fetched as
// 16-byte pairs (needs an even image width and a 16-byte aligned map).  A patch covered by one label (the
// common case inside a superpixel) costs one vote.  Otherwise labels are counted one distinct value per round,
// candidates taken in whatever order lanes hold them, ties resolved towards the smaller id (sppp.py:117-120).
// CENT: the same pass also accumulates, per image and label in [0, K), the pixel count and the sums of the x and y
// coordinates (models/sppp_mhla.py:226-262, the centroids of the dynamic positional encoding) into acc [B][3][K]
// (64-bit integers, exact and order independent) — the label map is then read ONCE per forward instead of twice.
// Only when the patches tile the whole image (img_h == img_w == grid * PS); the vote loop then runs until every
// pixel of a mixed patch has been counted instead of stopping as soon as the winner is known.
template <int PS, bool CENT>
__global__ void __launch_bounds__(256) sppp_dominant_vec_kernel(const int64_t* __restrict__ labels,
                                                                int64_t* __restrict__ dom, int B, int img_h,
                                                                int img_w, int grid, unsigned long long* __restrict__ acc,
                                                                int K) {
  constexpr int NV = PS * PS / 32;             // labels per lane
  constexpr int NP = NV / 2;                   // 16-byte pairs per lane
  constexpr int PAIRS_PER_ROW = PS / 2;        // lanes covering one patch row
  constexpr int ROWS_PER_LOAD = 32 / PAIRS_PER_ROW;
  const int P = grid * grid;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (int64_t)B * P) return;  // warp-uniform
  const int lane = threadIdx.x & 31;
  const int b = (int)(warp / P);
  const int p = (int)(warp - (int64_t)b * P);
  const int pi = p / grid, pj = p - pi * grid;
  const int64_t* src = labels + ((int64_t)b * img_h + (int64_t)pi * PS + lane / PAIRS_PER_ROW) * img_w +
                       (int64_t)pj * PS + 2 * (lane % PAIRS_PER_ROW);
  long long vals[NV];
#pragma unroll
  for (int k = 0; k < NP; ++k) {
    const longlong2 v = __ldcs(reinterpret_cast<const longlong2*>(src + (int64_t)k * ROWS_PER_LOAD * img_w));
    vals[2 * k] = v.x;
    vals[2 * k + 1] = v.y;
  }
  const long long v0 = __shfl_sync(0xffffffffu, vals[0], 0);
  bool differs = false;
#pragma unroll
  for (int t = 0; t < NV; ++t) differs |= vals[t] != v0;
  if (!__any_sync(0xffffffffu, differs)) {
    if (lane == 0) {
      dom[warp] = v0;
      if (CENT && v0 >= 0 && v0 < K) {
        unsigned long long* a = acc + (int64_t)b * 3 * K + v0;
        atomicAdd(a, (unsigned long long)(PS * PS));
        atomicAdd(a + K, (unsigned long long)(PS * (PS * pj * PS + PS * (PS - 1) / 2)));
        atomicAdd(a + 2 * K, (unsigned long long)(PS * (PS * pi * PS + PS * (PS - 1) / 2)));
      }
    }
    return;
  }
  unsigned alive = NV == 32 ? 0xffffffffu : ((1u << NV) - 1u);
  int remaining = PS * PS, best_cnt = 0;
  long long best_label = 0;
  const int x_lane = pj * PS + 2 * (lane % PAIRS_PER_ROW), y_lane = pi * PS + lane / PAIRS_PER_ROW;
  while (remaining > 0 && (CENT || remaining >= best_cnt)) {
    const unsigned has = __ballot_sync(0xffffffffu, alive != 0u);
    long long mine = 0;
#pragma unroll
    for (int t = NV - 1; t >= 0; --t)
      if ((alive >> t) & 1u) mine = vals[t];
    const long long cand = __shfl_sync(0xffffffffu, mine, __ffs(has) - 1);
    int c = 0, sx = 0, sy = 0;
#pragma unroll
    for (int t = 0; t < NV; ++t)
      if (((alive >> t) & 1u) && vals[t] == cand) {
        ++c;
        alive &= ~(1u << t);
        if (CENT) {
          sx += x_lane + (t & 1);
          sy += y_lane + (t >> 1) * ROWS_PER_LOAD;
        }
      }
    c = __reduce_add_sync(0xffffffffu, c);
    if (CENT) {
      sx = __reduce_add_sync(0xffffffffu, sx);
      sy = __reduce_add_sync(0xffffffffu, sy);
      if (lane == 0 && cand >= 0 && cand < K) {
        unsigned long long* a = acc + (int64_t)b * 3 * K + cand;
        atomicAdd(a, (unsigned long long)c);
        atomicAdd(a + K, (unsigned long long)sx);
        atomicAdd(a + 2 * K, (unsigned long long)sy);
      }
    }
    if (c > best_cnt || (c == best_cnt && cand < best_label)) {
      best_cnt = c;
      best_label = cand;
    }
    remaining -= c;
  }
  if (lane == 0) dom[warp] = best_label;
}

// ---- slots: one CTA per image, everything in shared memory ----------------------------------------------
// first[i]   = first run whose label equals that of run i    (run i is a "leader" when first[i] == i)
// slot(leader) = number of leaders before it               (= dict insertion order, sppp.py:123-126)
// slot[p]    = slot(first[run(p)]);  counts by shared-memory integer atomics;  offsets = exclusive scan of counts
// order      = patches grouped by slot, ascending inside a slot
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  __syncthreads();  // s_warp may still be read from the previous call
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    const int w = lane < nw ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += y;
    }
    if (lane < nw) s_warp[lane] = winc - w;
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[wid] + incl - v;
}

// Works on RUNS of equal consecutive dominant labels (superpixels are spatially coherent: ~g*sqrt(K) runs for g*g
// patches), because the first occurrence of a label always starts a run and every patch of a run shares its slot:
// the two quadratic searches (first occurrence, rank inside the slot) touch runs, not patches.
__global__ void __launch_bounds__(1024) sppp_slot_smem_kernel(const int64_t* __restrict__ dom,
                                                              int32_t* __restrict__ slot,
                                                              int32_t* __restrict__ num_slots,
                                                              int32_t* __restrict__ counts,
                                                              int64_t* __restrict__ slot_label,
                                                              int32_t* __restrict__ offsets,
                                                              int32_t* __restrict__ order, int P, int r_cap) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int Pp = (P + 3) & ~3;  // the 4-wide scan may read up to 3 entries past the last run (never taken: min(i, .))
  long long* s_hlab = reinterpret_cast<long long*>(s_raw);  // [Pp]  label of each run
  int* s_run = reinterpret_cast<int*>(s_hlab + Pp);         // [P]   run of each patch
  int* s_head = s_run + P;                                  // [P+1] first patch of each run, then P
  int* s_first = s_head + P + 1;                            // [P]   first run with the same label
  int* s_rslot = s_first + P;                               // [P]   slot of each run
  int* s_rrank = s_rslot + P;                               // [P]   patches of the same slot in earlier runs
  int* s_cnt = s_rrank + P;                                 // [P]   count per slot, then exclusive offsets
  int* s_warp = s_cnt + P;                                  // [33]  scan scratch
  const int b = blockIdx.x, tid = threadIdx.x, T = blockDim.x;
  const int64_t* d = dom + (int64_t)b * P;
  // 1. runs
  int base = 0;
  for (int p0 = 0; p0 < P; p0 += T) {
    const int p = p0 + tid;
    long long mine = 0;
    int head = 0;
    if (p < P) {
      mine = d[p];
      head = (p == 0 || d[p - 1] != mine) ? 1 : 0;
      s_cnt[p] = 0;
    }
    int total;
    const int ex = block_excl_scan(head, s_warp, &total);
    if (p < P) {
      const int ri = base + ex + head - 1;
      s_run[p] = ri;
      if (head) {
        s_head[ri] = p;
        s_hlab[ri] = mine;
      }
    }
    base += total;
  }
  const int Hn = base;
  if (tid == 0) s_head[Hn] = P;
  if (tid < Pp - Hn && tid < 4) s_hlab[Hn + tid] = 0;
  __syncthreads();
  // 2. first run with the same label (16-byte shared loads; lanes of a warp read the same address: broadcast)
  for (int i = tid; i < Hn; i += T) {
    const long long mine = s_hlab[i];
    int first = i;
    for (int q = 0; q < i; q += 4) {
      const longlong2 a = *reinterpret_cast<const longlong2*>(s_hlab + q);
      const longlong2 c = *reinterpret_cast<const longlong2*>(s_hlab + q + 2);
      const unsigned m = (a.x == mine ? 1u : 0u) | (a.y == mine ? 2u : 0u) | (c.x == mine ? 4u : 0u) |
                         (c.y == mine ? 8u : 0u);
      if (m) {
        first = min(i, q + __ffs(m) - 1);
        break;
      }
    }
    s_first[i] = first;
  }
  __syncthreads();
  // 3. leader runs (first[i] == i) take consecutive slots in patch order = the dict's insertion order
  base = 0;
  for (int i0 = 0; i0 < Hn; i0 += T) {
    const int i = i0 + tid;
    const int lead = (i < Hn && s_first[i] == i) ? 1 : 0;
    int total;
    const int ex = block_excl_scan(lead, s_warp, &total);
    if (lead) {
      const int r = base + ex;
      s_rslot[i] = r;
      if (r < r_cap) slot_label[(int64_t)b * r_cap + r] = s_hlab[i];
    }
    base += total;
  }
  const int R = base;
  __syncthreads();
  // 4. every run: its leader's slot, its share of the count, and the patches of its slot that come before it
  for (int i = tid; i < Hn; i += T) {
    const int f = s_first[i];
    const int r = s_rslot[f];
    if (f != i) s_rslot[i] = r;
    atomicAdd(&s_cnt[r], s_head[i + 1] - s_head[i]);
    int rank = 0;
    for (int q = f; q < i; ++q) rank += s_first[q] == f ? s_head[q + 1] - s_head[q] : 0;
    s_rrank[i] = rank;
  }
  __syncthreads();
  // 5. counts -> offsets
  base = 0;
  for (int r0 = 0; r0 < R; r0 += T) {
    const int r = r0 + tid;
    const int c = r < R ? s_cnt[r] : 0;
    int total;
    const int ex = block_excl_scan(c, s_warp, &total);
    if (r < R) {
      s_cnt[r] = base + ex;
      if (r <= r_cap) offsets[(int64_t)b * (r_cap + 1) + r] = base + ex;  // r == r_cap: end of the last kept row
      if (r < r_cap) counts[(int64_t)b * r_cap + r] = c;
    }
    base += total;
  }
  __syncthreads();
  if (tid == 0) num_slots[b] = R;
  for (int q = R + tid; q <= r_cap; q += T) {  // unused rows: offsets = P, counts = 0
    offsets[(int64_t)b * (r_cap + 1) + q] = P;
    if (q < r_cap) {
      counts[(int64_t)b * r_cap + q] = 0;
      slot_label[(int64_t)b * r_cap + q] = 0;
    }
  }
  // 6. patches: slot id and CSR position (runs of one slot are laid out in run order, patches ascending)
  int32_t* ord = order + (int64_t)b * P;
  for (int p = tid; p < P; p += T) {
    const int i = s_run[p];
    const int r = s_rslot[i];
    slot[(int64_t)b * P + p] = r;
    ord[s_cnt[r] + s_rrank[i] + p - s_head[i]] = p;
  }
}

// ---- superpixel centroids -----------------------------------------------------------------------------------
// models/sppp_mhla.py:226-262: for every image and every label s in [0, K) the mean of x / W and y / H over the
// pixels carrying s ((0.5, 0.5) for a label without pixels).  The reference loops over images and labels with a
// device sync per `if mask.sum() > 0`; a batched index_add_ is atomics-bound (1.2 ms per call at 32 x 512 x 512).
// Here each thread walks 8 consecutive pixels of one row (two-label 16-byte loads) and flushes a (count, sum x, y * count)
// triple per run of equal labels into the CTA's integer accumulators in shared memory (native integer atomics, exact and
// order-independent), CTAs add their partial sums to 64-bit global accumulators, a second tiny kernel divides.
constexpr int kCentPix = 8;  // pixels per thread
__global__ void __launch_bounds__(256) sppp_centroid_acc_kernel(const int64_t* __restrict__ labels, int B, int img_h,
                                                                int img_w, int K, int rows_per_cta, int chunks,
                                                                unsigned long long* __restrict__ acc) {
  extern __shared__ unsigned int s_cent[];  // [3][K]: count, sum x, sum y (< 2^32 per CTA: <= 32768 pixels, coordinates < 65536)
  const int b = blockIdx.x / chunks, ch = blockIdx.x - b * chunks;
  for (int i = threadIdx.x; i < 3 * K; i += blockDim.x) s_cent[i] = 0;
  __syncthreads();
  const int y0 = ch * rows_per_cta, y1 = min(img_h, y0 + rows_per_cta);
  const int groups = (img_w + kCentPix - 1) / kCentPix;  // 8-pixel groups per row
  const bool vec = (img_w % 2 == 0) && ((uintptr_t)labels % 16 == 0);
  for (int i = threadIdx.x; i < (y1 - y0) * groups; i += blockDim.x) {
    const int y = y0 + i / groups, x0 = (i % groups) * kCentPix;
    const int64_t* src = labels + ((int64_t)b * img_h + y) * img_w + x0;
    long long v[kCentPix];
    if (vec && x0 + kCentPix <= img_w) {
#pragma unroll
      for (int k = 0; k < kCentPix / 2; ++k) {
        const longlong2 t = __ldcs(reinterpret_cast<const longlong2*>(src) + k);
        v[2 * k] = t.x;
        v[2 * k + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int k = 0; k < kCentPix; ++k) v[k] = (x0 + k < img_w) ? src[k] : -1;
    }
    long long cur = v[0];
    int n = 0, sx = 0;
#pragma unroll
    for (int k = 0; k <= kCentPix; ++k) {
      const long long l = k < kCentPix ? v[k] : cur - 1;  // sentinel: flush the last run
      if (k == kCentPix || l != cur) {
        if (n > 0 && cur >= 0 && cur < K) {
          atomicAdd(&s_cent[(int)cur], (unsigned)n);
          atomicAdd(&s_cent[K + (int)cur], (unsigned)sx);
          atomicAdd(&s_cent[2 * K + (int)cur], (unsigned)(n * y));
        }
        cur = l;
        n = 0;
        sx = 0;
      }
      if (k < kCentPix) {
        ++n;
        sx += x0 + k;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * K; i += blockDim.x)
    if (s_cent[i]) atomicAdd(acc + (int64_t)b * 3 * K + i, (unsigned long long)s_cent[i]);
}

__global__ void __launch_bounds__(256) sppp_centroid_fin_kernel(const unsigned long long* __restrict__ acc, int B, int K,
                                                                int img_h, int img_w, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * K) return;
  const int b = i / K, k = i - b * K;
  const unsigned long long n = acc[(int64_t)b * 3 * K + k];
  float cx = 0.5f, cy = 0.5f;
  if (n > 0) {
    cx = (float)((double)acc[(int64_t)b * 3 * K + K + k] / (double)n / (double)img_w);
    cy = (float)((double)acc[(int64_t)b * 3 * K + 2 * K + k] / (double)n / (double)img_h);
  }
  out[2 * i] = cx;
  out[2 * i + 1] = cy;
}


// ---- pool forward: TMA-staged tiles ------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  bind_context();
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

constexpr int kSliceBytes = 128;   // bytes of one patch row per tile
constexpr int kPoolThreads = 256;  // 16 slot lanes x 16 column groups of 8 bytes

template <typename TIn> struct Group8;  // the 8 bytes one thread owns of a tile row
template <> struct Group8<__nv_bfloat16> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void add(const unsigned char* p, float (&acc)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    acc[0] += __uint_as_float(u.x << 16);
    acc[1] += __uint_as_float(u.x & 0xffff0000u);
    acc[2] += __uint_as_float(u.y << 16);
    acc[3] += __uint_as_float(u.y & 0xffff0000u);
  }
};
template <> struct Group8<float> {
  static constexpr int N = 2;
  static __device__ __forceinline__ void add(const unsigned char* p, float (&acc)[2]) {
    const float2 u = *reinterpret_cast<const float2*>(p);
    acc[0] += u.x;
    acc[1] += u.y;
  }
};
template <int N> __device__ __forceinline__ void store_group(float* o, const float (&v)[N]) {
  if constexpr (N == 4) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
  else *reinterpret_cast<float2*>(o) = make_float2(v[0], v[1]);
}
template <int N> __device__ __forceinline__ void store_group(__nv_bfloat16* o, const float (&v)[N]) {
  if constexpr (N == 4) *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  else *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(v[0], v[1]);
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ptx::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ptx::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void lds_group(const float* p, float (&v)[N]) {
  if constexpr (N == 4) {
    const float4 u = *reinterpret_cast<const float4*>(p);
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
  } else {
    const float2 u = *reinterpret_cast<const float2*>(p);
    v[0] = u.x; v[1] = u.y;
  }
}
template <int N> __device__ __forceinline__ void sts_group(float* p, const float (&v)[N]) {
  if constexpr (N == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}

// (a) P <= 256: the whole [P x 128 B] tile of an item = (image b, 128-byte column slice s) is one TMA box.
// Persistent CTAs; a ring of `stages` boxes runs ahead across item boundaries, and the CSR of the NEXT item is
// fetched by cp.async into the other of two buffers while this tile is summed, so nothing but the barrier wait
// sits between two tiles.  Tensor map: x viewed as [B*P rows, D cols], box = P x (128 / sizeof(TIn)) columns.
// Thread (lane16 = tid / 16, cg = tid % 16) sums the 8-byte column group cg over the CSR lists of slots lane16,
// lane16 + 16, ... in ascending patch order.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kPoolThreads) sppp_pool_fwd_tma_kernel(
    const __grid_constant__ CUtensorMap tm, const int32_t* __restrict__ order, const int32_t* __restrict__ offsets,
    const int32_t* __restrict__ num_slots, TOut* __restrict__ out, int P, int R, int D, int r_cap, int nslices,
    int nitems, int stages) {
  constexpr int NV = Group8<TIn>::N;
  constexpr int CW = kSliceBytes / (int)sizeof(TIn);
  // All shared memory is dynamic so that the TMA destination sits at the (128-byte aligned) base of the window.
  extern __shared__ __align__(128) unsigned char s_tile0[];
  const uint32_t tile_bytes = (uint32_t)P * kSliceBytes;
  unsigned char* ring = s_tile0;
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(ring + (size_t)stages * tile_bytes);  // [4]
  int* s_order = reinterpret_cast<int*>(s_bar + 4);  // [2][P]
  int* s_off = s_order + 2 * P;                      // [2][R+1]
  const int tid = threadIdx.x;
  const int cg = tid & 15, sl = tid >> 4;
  const int my_items = (nitems - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  auto issue = [&](int j) {  // one thread
    const int item = blockIdx.x + j * gridDim.x;
    const int b = item / nslices, s = item - b * nslices;
    const int st = j % stages;
    const uint32_t bar = ptx::smem_u32(&s_bar[st]);
    ptx::mbar_expect_tx(bar, tile_bytes);
    ptx::tma_load_2d(ptx::smem_u32(ring + (size_t)st * tile_bytes), &tm, bar, s * CW, b * P);
  };
  auto fetch_csr = [&](int j, int buf) {  // all threads, asynchronous
    const int b = (blockIdx.x + j * gridDim.x) / nslices;
    int* so = s_order + buf * P;
    int* sf = s_off + buf * (R + 1);
    for (int i = tid; i < P; i += kPoolThreads) cp_async4(so + i, order + (int64_t)b * P + i);
    for (int r = tid; r <= R && r <= r_cap; r += kPoolThreads) cp_async4(sf + r, offsets + (int64_t)b * (r_cap + 1) + r);
  };
  pdl_launch_dependents();
  if (tid == 0) {
    ptx::prefetch_tmap(&tm);
    for (int i = 0; i < stages; ++i) ptx::mbar_init(ptx::smem_u32(&s_bar[i]), 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  pdl_wait();   // everything above is CTA-local set-up; from here on global memory is touched
  if (tid == 0)
    for (int j = 0; j < stages && j < my_items; ++j) issue(j);
  fetch_csr(0, 0);
  int ns_next = min(min(num_slots[blockIdx.x / nslices], r_cap), R);

  for (int j = 0; j < my_items; ++j) {
    const int item = blockIdx.x + j * gridDim.x;
    const int b = item / nslices, s = item - b * nslices;
    const int ns = ns_next;
    const int buf = j & 1;
    const int* so = s_order + buf * P;
    const int* sf = s_off + buf * (R + 1);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();  // this item's CSR is in place (and the barriers are initialised, first time round)
    if (j + 1 < my_items) {
      ns_next = min(min(num_slots[(blockIdx.x + (j + 1) * gridDim.x) / nslices], r_cap), R);
      fetch_csr(j + 1, buf ^ 1);  // that buffer's readers finished before the barrier that ended item j-1
    }
    const int st = j % stages;
    ptx::mbar_wait(ptx::smem_u32(&s_bar[st]), (uint32_t)(j / stages) & 1u);
    const unsigned char* tile = ring + (size_t)st * tile_bytes + cg * 8;
    const int d0 = s * CW + cg * NV;
    for (int r = sl; r < R; r += 16) {
      const int beg = r < ns ? sf[r] : 0, end = r < ns ? sf[r + 1] : 0;
      float acc[NV];
#pragma unroll
      for (int e = 0; e < NV; ++e) acc[e] = 0.f;
#pragma unroll 4
      for (int t = beg; t < end; ++t) Group8<TIn>::add(tile + (size_t)so[t] * kSliceBytes, acc);
      if (d0 < D) {
        const float n = (float)max(end - beg, 1);
#pragma unroll
        for (int e = 0; e < NV; ++e) acc[e] = acc[e] / n;
        store_group<NV>(out + ((int64_t)b * R + r) * D + d0, acc);
      }
    }
    __syncthreads();  // every thread is done with this stage and this CSR buffer
    if (tid == 0 && j + stages < my_items) issue(j + stages);
  }
}

// (b) P > 256: rows are streamed in CSR ORDER (slot-major, ascending patch id inside a slot), so the shared-memory
// tile is already sorted by slot and the pooling is a segmented sum over a contiguous stream.  Per chunk of 128
// CSR positions each thread copies 4 x 16 bytes with cp.async (3-stage ring), then
//   phase 1: thread (lane16, cg) sums the 8-byte column group cg of positions [8*lane16, 8*lane16 + 8); a slot that
//            begins and ends inside the lane is finished on the spot; the part of a slot that began in an earlier
//            lane is published as the lane's HEAD; the part of a slot that goes on past the lane is its TAIL;
//   phase 2: the thread holding the tail of a slot adds the heads of the following lanes until the slot ends (1-2
//            for compact superpixels) and writes the mean; a slot still open at the end of the chunk is carried.
// Every position is touched once, the work per lane is equal whatever the shape of the superpixels, and the order
// of the additions is fixed.
constexpr int kSortedThreads = 256;
constexpr int kRowsPerLane = 8;
constexpr int kSortedChunkBytes = 16 * 1024;  // per stage: 128 positions of a 128-byte slice, or 256 of a 64-byte one

// SLICE = bytes of a patch row one item covers: 128, or 64 when 128 would leave SMs without an item (a CTA's rate is
// bound by its own wait -> sum -> wait chain, so small batches want more, narrower items).
template <typename TIn, typename TOut, int SLICE>
__global__ void __launch_bounds__(kSortedThreads) sppp_pool_fwd_sorted_kernel(
    const TIn* __restrict__ x, const int32_t* __restrict__ order, const int32_t* __restrict__ offsets,
    const int32_t* __restrict__ num_slots, TOut* __restrict__ out, int P, int R, int D, int r_cap, int nslices,
    int nitems) {
  constexpr int kSortedStages = 3;
  constexpr int NV = Group8<TIn>::N;
  constexpr int CW = SLICE / (int)sizeof(TIn);
  constexpr int CGS = SLICE / 8;                      // 8-byte column groups per row
  constexpr int kLanes = kSortedThreads / CGS;        // row lanes
  constexpr int kChunkPos = kLanes * kRowsPerLane;    // CSR positions per chunk
  constexpr int PIECES = SLICE / 16;                  // 16-byte copies per row
  constexpr int kCopyRows = kSortedThreads / PIECES;  // rows per copy pass
  constexpr uint32_t kChunkBytes = kChunkPos * SLICE;
  static_assert(kChunkBytes == kSortedChunkBytes, "stage size");
  extern __shared__ __align__(128) unsigned char s_tile1[];
  unsigned char* ring = s_tile1;                                                         // [3][128][128 B]
  float* s_head = reinterpret_cast<float*>(ring + kSortedStages * kChunkBytes);        // [kLanes][CGS][NV]
  float* s_carry = s_head + kLanes * CGS * NV;                                              // [2][16][NV]
  int* s_cont = reinterpret_cast<int*>(s_carry + 2 * CGS * NV);                        // [kLanes] head goes on past its lane
  int* s_off = s_cont + kLanes;                                                            // [R+1]
  int* s_order = s_off + (R + 1);                                                      // [P]
  short* s_lslot = reinterpret_cast<short*>(s_order + P);                              // [P/8 + 1] slot holding position 8*i
  const int tid = threadIdx.x;
  const int cg = tid % CGS, sl = tid / CGS;
  const int crow = tid / PIECES, cpiece = tid % PIECES;  // copy role: position crow (+kCopyRows per pass), piece cpiece

  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int b = item / nslices, s = item - b * nslices;
    const int ns = min(min(num_slots[b], r_cap), R);
    __syncthreads();  // the previous item's readers are done with the CSR buffers
    for (int i = tid; i < P; i += kSortedThreads) cp_async4(s_order + i, order + (int64_t)b * P + i);
    for (int r = tid; r <= ns; r += kSortedThreads) cp_async4(s_off + r, offsets + (int64_t)b * (r_cap + 1) + r);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    const int covered = s_off[ns];  // positions [0, covered) belong to slots [0, ns)
    const int nchunks = (covered + kChunkPos - 1) / kChunkPos;
    for (int i = tid; i * kRowsPerLane < covered; i += kSortedThreads) {
      const int t = i * kRowsPerLane;
      int lo = 0, hi = ns;  // last r with s_off[r] <= t
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= t) lo = mid;
        else hi = mid;
      }
      s_lslot[i] = (short)lo;
    }
    const int d0 = s * CW + cg * NV;
    // slots without patches (or beyond r_cap) pool to zero
    for (int r = sl; r < R; r += kLanes) {
      if ((r >= ns || s_off[r + 1] == s_off[r]) && d0 < D) {
        float z[NV];
#pragma unroll
        for (int e = 0; e < NV; ++e) z[e] = 0.f;
        store_group<NV>(out + ((int64_t)b * R + r) * D + d0, z);
      }
    }
    // copy role: 16-byte piece `cpiece` of positions crow, crow + 32, ...; positions past `covered` re-read the last
    // row (never summed), so the loop carries no predicate; 32-bit row offsets (the host checked P*D*size < 2^31)
    const unsigned char* xs = reinterpret_cast<const unsigned char*>(x) +
                              ((int64_t)b * P * D + (int64_t)s * CW) * sizeof(TIn) + cpiece * 16;
    const bool piece_ok = (int64_t)s * SLICE + cpiece * 16 < (int64_t)D * (int64_t)sizeof(TIn);
    const uint32_t pitch = (uint32_t)D * (uint32_t)sizeof(TIn);
    const uint32_t dst0 = ptx::smem_u32(ring) + crow * SLICE + cpiece * 16;
    auto copy_chunk = [&](int c) {  // all threads
      if (c < nchunks && piece_ok) {
        const uint32_t dst = dst0 + (uint32_t)(c % kSortedStages) * kChunkBytes;
        const int t0 = c * kChunkPos + crow;
#pragma unroll
        for (int q = 0; q < kChunkPos / kCopyRows; ++q) {
          const int t = min(t0 + kCopyRows * q, covered - 1);
          const unsigned char* src = xs + (uint32_t)s_order[t] * pitch;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + q * kCopyRows * SLICE), "l"(src) : "memory");
        }
      }
      cp_async_commit();
    };
#pragma unroll
    for (int c = 0; c < kSortedStages - 1; ++c) copy_chunk(c);
    int carry_buf = 0;
    for (int c = 0; c < nchunks; ++c) {
      copy_chunk(c + kSortedStages - 1);  // its stage was released by the barrier that closed chunk c-1
      cp_async_wait<kSortedStages - 1>();
      __syncthreads();  // chunk c is in shared memory; phase 2 of chunk c-1 is over (s_head / s_cont are free)
      const unsigned char* tile = ring + (size_t)(c % kSortedStages) * kChunkBytes + cg * 8;
      // ---- phase 1
      const int a = c * kChunkPos + sl * kRowsPerLane;  // first position of this lane
      const int lane_end = min(a + kRowsPerLane, covered);
      float acc[NV], tail[NV];
#pragma unroll
      for (int e = 0; e < NV; ++e) acc[e] = 0.f;
      int tail_slot = -1;       // slot whose sum continues past this lane (this thread then owns its completion)
      bool is_head = false;     // the run being summed began before this lane
      if (a < covered) {
        int r = s_lslot[a / kRowsPerLane];
        int nxt = s_off[r + 1];
        is_head = s_off[r] < a;
#pragma unroll
        for (int i = 0; i < kRowsPerLane; ++i) {
          const int t = a + i;
          if (t < lane_end) {
            if (t >= nxt) {  // slot r ended at t: finish or publish it, move to the slot holding t
              if (is_head) {
                sts_group<NV>(s_head + (sl * CGS + cg) * NV, acc);
                if (cg == 0) s_cont[sl] = 0;
                is_head = false;
              } else if (d0 < D) {
                const float n = (float)(nxt - s_off[r]);
                float v[NV];
#pragma unroll
                for (int e = 0; e < NV; ++e) v[e] = acc[e] / n;
                store_group<NV>(out + ((int64_t)b * R + r) * D + d0, v);
              }
#pragma unroll
              for (int e = 0; e < NV; ++e) acc[e] = 0.f;
              do {
                ++r;
                nxt = s_off[r + 1];
              } while (t >= nxt);  // skips empty slots
            }
            Group8<TIn>::add(tile + (size_t)(sl * kRowsPerLane + i) * SLICE, acc);
          }
        }
        // the run open at the end of the lane
        const bool goes_on = nxt > lane_end;  // more positions of slot r follow (next lane or next chunk)
        if (is_head) {
          sts_group<NV>(s_head + (sl * CGS + cg) * NV, acc);
          if (cg == 0) s_cont[sl] = goes_on ? 1 : 0;
        } else if (goes_on) {
          tail_slot = r;
#pragma unroll
          for (int e = 0; e < NV; ++e) tail[e] = acc[e];
        } else if (d0 < D) {
          const float n = (float)(nxt - s_off[r]);
          float v[NV];
#pragma unroll
          for (int e = 0; e < NV; ++e) v[e] = acc[e] / n;
          store_group<NV>(out + ((int64_t)b * R + r) * D + d0, v);
        }
      }
      __syncthreads();
      // ---- phase 2: owners complete their slots.  Lane 0 also owns the slot carried in from the previous chunk.
      auto complete = [&](int r, float (&sum)[NV], int l2) {
        for (; l2 < kLanes; ++l2) {
          if (c * kChunkPos + l2 * kRowsPerLane >= covered) break;  // cannot happen for an open slot; safety
          float v[NV];
          lds_group<NV>(s_head + (l2 * CGS + cg) * NV, v);
#pragma unroll
          for (int e = 0; e < NV; ++e) sum[e] += v[e];
          if (!s_cont[l2]) {
            if (d0 < D) {
              const float n = (float)(s_off[r + 1] - s_off[r]);
#pragma unroll
              for (int e = 0; e < NV; ++e) sum[e] = sum[e] / n;
              store_group<NV>(out + ((int64_t)b * R + r) * D + d0, sum);
            }
            return;
          }
        }
        sts_group<NV>(s_carry + ((carry_buf ^ 1) * CGS + cg) * NV, sum);  // still open: carry into the next chunk
      };
      if (sl == 0 && c > 0) {
        const int a0 = c * kChunkPos;
        const int r0 = s_lslot[a0 / kRowsPerLane];
        if (s_off[r0] < a0) {  // the first slot of this chunk began in an earlier chunk
          float sum[NV];
          lds_group<NV>(s_carry + (carry_buf * CGS + cg) * NV, sum);
          complete(r0, sum, 0);
        }
      }
      if (tail_slot >= 0) complete(tail_slot, tail, sl + 1);
      carry_buf ^= 1;
    }
    cp_async_wait<0>();
  }
}

// ---- pool backward: pre-divided gradient rows in shared memory, 16-byte stores ----------------------------
constexpr int kBwdCols = 64;
constexpr int kBwdThreads = 128;  // small CTAs: the whole grid is resident at once, prologue latencies overlap
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kBwdThreads) sppp_pool_bwd_tile_kernel(const TIn* __restrict__ dout,
                                                                         const int32_t* __restrict__ slot,
                                                                         const int32_t* __restrict__ counts,
                                                                         TOut* __restrict__ dx, int P, int R, int D,
                                                                         int r_cap, int nslices) {
  constexpr int NVO = 16 / (int)sizeof(TOut);  // columns per 16-byte store
  constexpr int TPR = kBwdCols / NVO;          // threads per patch row
  constexpr int ROWS = kBwdThreads / TPR;
  extern __shared__ __align__(16) float s_g[];  // [R][64] gradient rows / count, then [P] slot ids
  int* s_slot = reinterpret_cast<int*>(s_g + (size_t)R * kBwdCols);
  const int tid = threadIdx.x;
  const int b = blockIdx.x / nslices, s = blockIdx.x - b * nslices;
  const int col0 = s * kBwdCols;
  for (int i = tid; i < P; i += kBwdThreads) {
    const int r = slot[(int64_t)b * P + i];
    s_slot[i] = (r >= 0 && r < R && r < r_cap) ? r : -1;
  }
  for (int i = tid; i < R * kBwdCols; i += kBwdThreads) {
    const int r = i / kBwdCols, c = col0 + (i % kBwdCols);
    float v = 0.f;
    if (c < D && r < r_cap)
      v = Elem<TIn>::ld(dout + ((int64_t)b * R + r) * D + c) / (float)max(counts[(int64_t)b * r_cap + r], 1);
    s_g[i] = v;
  }
  __syncthreads();
  const int cgo = tid % TPR, rl = tid / TPR;
  const int c = col0 + cgo * NVO;
  if (c >= D) return;
  TOut* o = dx + (int64_t)b * P * D + c;
#pragma unroll 4
  for (int p = rl; p < P; p += ROWS) {
    const int r = s_slot[p];
    float f[NVO];
    if (r >= 0) {
      const float4* g = reinterpret_cast<const float4*>(s_g + r * kBwdCols + cgo * NVO);
#pragma unroll
      for (int e = 0; e < NVO / 4; ++e) {
        const float4 v = g[e];
        f[4 * e] = v.x; f[4 * e + 1] = v.y; f[4 * e + 2] = v.z; f[4 * e + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < NVO; ++e) f[e] = 0.f;
    }
    if constexpr (NVO == 8) {
      store8(o + (int64_t)p * D, f);
    } else {
      *reinterpret_cast<float4*>(o + (int64_t)p * D) = make_float4(f[0], f[1], f[2], f[3]);
    }
  }
}

// Whole-row variant: one CTA per (image, chunk of consecutive patch rows).  The R gradient rows of the image, already
// divided by their counts and rounded to the OUTPUT type (so that replicating them is a pure copy and gives the same
// bits as dividing per patch), sit in shared memory; the CTA then writes its rows x D block of dx — one contiguous
// range of global memory — 16 bytes per thread, consecutive threads at consecutive addresses.  Column-sliced CTAs
// (above) write 128-byte pieces D*es bytes apart, which costs DRAM write efficiency: 0.60 of the copy bandwidth at
// best; contiguous rows stream.
constexpr int kBwdRowsThreads = 256;
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kBwdRowsThreads) sppp_pool_bwd_rows_kernel(const TIn* __restrict__ dout,
                                                                             const int32_t* __restrict__ slot,
                                                                             const int32_t* __restrict__ counts,
                                                                             TOut* __restrict__ dx, int P, int R, int D,
                                                                             int r_cap, int nchunks, int rows_per_chunk) {
  constexpr int NVO = 16 / (int)sizeof(TOut);
  extern __shared__ __align__(16) unsigned char s_rows_raw[];
  TOut* s_g = reinterpret_cast<TOut*>(s_rows_raw);                       // [R][D] in the output type
  int* s_slot = reinterpret_cast<int*>(s_rows_raw + (size_t)R * D * sizeof(TOut));   // [rows_per_chunk]
  const int tid = threadIdx.x;
  const int b = blockIdx.x / nchunks, ch = blockIdx.x - b * nchunks;
  const int p0 = ch * rows_per_chunk, p1 = min(P, p0 + rows_per_chunk);
  pdl_launch_dependents();
  pdl_wait();
  for (int i = tid; i < p1 - p0; i += kBwdRowsThreads) {
    const int r = slot[(int64_t)b * P + p0 + i];
    s_slot[i] = (r >= 0 && r < R && r < r_cap) ? r : -1;
  }
  const int cpr = D / NVO;  // 16-byte chunks per row
  for (int i = tid; i < R * cpr; i += kBwdRowsThreads) {
    const int r = i / cpr, c = (i - r * cpr) * NVO;
    float f[NVO];
    if (r < r_cap) {
      const float n = (float)max(counts[(int64_t)b * r_cap + r], 1);
      const TIn* src = dout + ((int64_t)b * R + r) * D + c;
      if constexpr (NVO == 8) {
        load8(src, f);
      } else {
#pragma unroll
        for (int e = 0; e < NVO; ++e) f[e] = Elem<TIn>::ld(src + e);
      }
#pragma unroll
      for (int e = 0; e < NVO; ++e) f[e] = f[e] / n;
    } else {
#pragma unroll
      for (int e = 0; e < NVO; ++e) f[e] = 0.f;
    }
    if constexpr (NVO == 8) store8(s_g + (size_t)r * D + c, f);
    else *reinterpret_cast<float4*>(s_g + (size_t)r * D + c) = make_float4(f[0], f[1], f[2], f[3]);
  }
  __syncthreads();
  uint4* o = reinterpret_cast<uint4*>(dx + ((int64_t)b * P + p0) * D);
  const int total = (p1 - p0) * cpr;
  int row = tid / cpr, cc = tid - row * cpr;            // kBwdRowsThreads / cpr and % cpr advance (row, cc) per step
  const int drow = kBwdRowsThreads / cpr, dcc = kBwdRowsThreads - drow * cpr;
#pragma unroll 4
  for (int i = tid; i < total; i += kBwdRowsThreads) {
    const int r = s_slot[row];
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r >= 0) v = *reinterpret_cast<const uint4*>(s_g + (size_t)r * D + cc * NVO);
    o[i] = v;
    row += drow;
    cc += dcc;
    if (cc >= cpr) { cc -= cpr; ++row; }
  }
}

// =====================================================================================================
// 'max' and 'attention' pooling (models/sppp.py:178-184 / 211-216): the non-default SuperpixelPooling variants.
// One 128-thread CTA per (image, slot) walks the slot's CSR list in ascending patch order; threads own columns.
// These are completeness kernels (no reference config selects them): coalesced and deterministic, not tuned.
// =====================================================================================================
constexpr int kVarThreads = 128;

__device__ __forceinline__ float block_sum_128(float v, float* s_red) {  // every thread gets the sum; s_red [4]
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  return s_red[0] + s_red[1] + s_red[2] + s_red[3];
}

// out[b,r,c] = max over the slot's patches of x[b,p,c]; argmax[b,r,c] = that patch (first one on ties), -1 if none.
template <typename TIn>
__global__ void __launch_bounds__(kVarThreads) sppp_pool_max_fwd_kernel(const TIn* __restrict__ x,
                                                                        const int32_t* __restrict__ order,
                                                                        const int32_t* __restrict__ offsets,
                                                                        const int32_t* __restrict__ num_slots,
                                                                        float* __restrict__ out,
                                                                        int32_t* __restrict__ argmax, int P, int R, int D,
                                                                        int r_cap) {
  const int b = blockIdx.x / R, r = blockIdx.x - b * R;
  int beg = 0, end = 0;
  if (r < r_cap && r < num_slots[b]) {
    beg = offsets[(int64_t)b * (r_cap + 1) + r];
    end = offsets[(int64_t)b * (r_cap + 1) + r + 1];
  }
  const int32_t* ord = order + (int64_t)b * P;
  const TIn* xb = x + (int64_t)b * P * D;
  for (int c = threadIdx.x; c < D; c += kVarThreads) {
    float best = 0.f;
    int arg = -1;
    for (int t = beg; t < end; ++t) {
      const int p = ord[t];
      const float v = Elem<TIn>::ld(xb + (int64_t)p * D + c);
      if (arg < 0 || v > best) {
        best = v;
        arg = p;
      }
    }
    out[((int64_t)b * R + r) * D + c] = best;
    argmax[((int64_t)b * R + r) * D + c] = arg;
  }
}

// dx (zeroed by the caller) [b, argmax[b,r,c], c] = dout[b,r,c]; slots are disjoint, so no two threads meet
template <typename TOut>
__global__ void __launch_bounds__(256) sppp_pool_max_bwd_kernel(const float* __restrict__ dout,
                                                                const int32_t* __restrict__ argmax,
                                                                TOut* __restrict__ dx, int64_t total, int P, int R, int D) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int p = argmax[i];
  if (p < 0) return;
  const int c = (int)(i % D);
  const int64_t b = i / ((int64_t)R * D);
  Elem<TOut>::st(dx + (b * P + p) * D + c, dout[i]);
}

// weights[b,p] = softmax over the slot's patches of sum_c x[b,p,c]; out[b,r,:] = sum_p weights[b,p] x[b,p,:]
template <typename TIn>
__global__ void __launch_bounds__(kVarThreads) sppp_pool_attn_fwd_kernel(const TIn* __restrict__ x,
                                                                         const int32_t* __restrict__ order,
                                                                         const int32_t* __restrict__ offsets,
                                                                         const int32_t* __restrict__ num_slots,
                                                                         float* __restrict__ out,
                                                                         float* __restrict__ weights, int P, int R, int D,
                                                                         int r_cap) {
  extern __shared__ float s_w[];  // [slot length] logits, then weights
  __shared__ float s_red[4];
  const int b = blockIdx.x / R, r = blockIdx.x - b * R;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int beg = 0, end = 0;
  if (r < r_cap && r < num_slots[b]) {
    beg = offsets[(int64_t)b * (r_cap + 1) + r];
    end = offsets[(int64_t)b * (r_cap + 1) + r + 1];
  }
  const int n = end - beg;
  const int32_t* ord = order + (int64_t)b * P + beg;
  const TIn* xb = x + (int64_t)b * P * D;
  // logits: one warp per patch, lane-strided partial sums, shuffle tree (fixed order)
  for (int k = wid; k < n; k += kVarThreads / 32) {
    const TIn* row = xb + (int64_t)ord[k] * D;
    float sum = 0.f;
    for (int c = lane; c < D; c += 32) sum += Elem<TIn>::ld(row + c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) s_w[k] = sum;
  }
  __syncthreads();
  float mx = -CUDART_INF_F;
  for (int k = threadIdx.x; k < n; k += kVarThreads) mx = fmaxf(mx, s_w[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) s_red[wid] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  float part = 0.f;
  for (int k = threadIdx.x; k < n; k += kVarThreads) {
    const float e = expf(s_w[k] - mx);
    s_w[k] = e;
    part += e;
  }
  const float denom = block_sum_128(part, s_red);  // barriers inside: s_w is complete afterwards
  const float inv = n > 0 ? 1.f / denom : 0.f;
  for (int k = threadIdx.x; k < n; k += kVarThreads) {
    const float w = s_w[k] * inv;
    s_w[k] = w;
    weights[(int64_t)b * P + ord[k]] = w;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += kVarThreads) {
    float acc = 0.f;
    for (int k = 0; k < n; ++k) acc = fmaf(s_w[k], Elem<TIn>::ld(xb + (int64_t)ord[k] * D + c), acc);
    out[((int64_t)b * R + r) * D + c] = acc;
  }
}

// dx[p,c] = w_p dy[c] + w_p (dy . x_p - dy . y_r)   (dx zeroed by the caller for patches outside every kept slot)
template <typename TIn>
__global__ void __launch_bounds__(kVarThreads) sppp_pool_attn_bwd_kernel(const TIn* __restrict__ x,
                                                                         const float* __restrict__ dout,
                                                                         const float* __restrict__ out,
                                                                         const float* __restrict__ weights,
                                                                         const int32_t* __restrict__ order,
                                                                         const int32_t* __restrict__ offsets,
                                                                         const int32_t* __restrict__ num_slots,
                                                                         TIn* __restrict__ dx, int P, int R, int D,
                                                                         int r_cap) {
  __shared__ float s_red[4];
  const int b = blockIdx.x / R, r = blockIdx.x - b * R;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int beg = 0, end = 0;
  if (r < r_cap && r < num_slots[b]) {
    beg = offsets[(int64_t)b * (r_cap + 1) + r];
    end = offsets[(int64_t)b * (r_cap + 1) + r + 1];
  }
  const int n = end - beg;
  const int32_t* ord = order + (int64_t)b * P + beg;
  const float* dy = dout + ((int64_t)b * R + r) * D;
  const float* y = out + ((int64_t)b * R + r) * D;
  float part = 0.f;
  for (int c = threadIdx.x; c < D; c += kVarThreads) part = fmaf(dy[c], y[c], part);
  const float t_r = block_sum_128(part, s_red);
  for (int k = wid; k < n; k += kVarThreads / 32) {
    const int p = ord[k];
    const TIn* row = x + ((int64_t)b * P + p) * D;
    float dot = 0.f;
    for (int c = lane; c < D; c += 32) dot = fmaf(dy[c], Elem<TIn>::ld(row + c), dot);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float w = weights[(int64_t)b * P + p];
    const float dl = w * (dot - t_r);
    TIn* drow = dx + ((int64_t)b * P + p) * D;
    for (int c = lane; c < D; c += 32) Elem<TIn>::st(drow + c, fmaf(w, dy[c], dl));
  }
}

// Tiled path when the rows can be described to TMA (16-byte aligned base and row pitch) and the tile ring fits.
template <typename TIn, typename TOut>
int launch_pool_fwd(const void* x, const int32_t* order, const int32_t* offsets, const int32_t* num_slots,
                    void* out, int B, int P, int R, int D, int r_cap, cudaStream_t st) {
  constexpr int CW = kSliceBytes / (int)sizeof(TIn);
  constexpr int NV = Group8<TIn>::N;
  const int nslices = ceil_div(D, CW);
  const int nitems = B * nslices;
  const bool aligned = ((uintptr_t)x % 16 == 0) && (((size_t)D * sizeof(TIn)) % 16 == 0) &&
                       ((uintptr_t)out % 16 == 0) && (D % 4 == 0) && R < 32768 && (int64_t)B * P < INT_MAX &&
                       (int64_t)B * nslices < INT_MAX;
  if (aligned && P <= 256 && encode_fn() != nullptr) {  // (a) one TMA box per item
    const size_t tile_bytes = (size_t)P * kSliceBytes;
    const int stages = tile_bytes > 12 * 1024 ? 2 : 4;
    const size_t smem = stages * tile_bytes + 32 + 2 * ((size_t)P + R + 1) * 4;
    if (smem <= 200 * 1024) {
      CUtensorMap tm;
      cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)B * P};
      cuuint64_t gstr[1] = {(cuuint64_t)D * sizeof(TIn)};
      cuuint32_t box[2] = {(cuuint32_t)CW, (cuuint32_t)P};
      cuuint32_t estr[2] = {1, 1};
      const CUresult r = encode_fn()(
          &tm, sizeof(TIn) == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
          const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("sppp_pool_fwd: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
        return FAVIT_ERR_CUDA;
      }
      static bool configured = false;  // per instantiation
      if (!configured) {
        FAVIT_CHECK_CUDA(cudaFuncSetAttribute(sppp_pool_fwd_tma_kernel<TIn, TOut>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
      }
      int per_sm = (int)((220 * 1024) / (smem + 1024));
      per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
      const int grid = nitems < num_sms() * per_sm ? nitems : num_sms() * per_sm;
      note_kernel("sppp_pool_fwd_tma_kernel grid=%d items=%d stages=%d", grid, nitems, stages);
      FAVIT_CHECK_CUDA(launch_pdl(sppp_pool_fwd_tma_kernel<TIn, TOut>, dim3(grid), dim3(kPoolThreads), smem, st, tm, order,
                                  offsets, num_slots, (TOut*)out, P, R, D, r_cap, nslices, nitems, stages));
      FAVIT_CHECK_LAUNCH();
      return FAVIT_OK;
    }
  }
  if (aligned && P > 256 && (int64_t)P * D * (int64_t)sizeof(TIn) < INT_MAX) {  // (b) rows streamed in CSR order
    const bool narrow = nitems < 2 * num_sms();  // 64-byte slices: twice the items
    const int slice = narrow ? 64 : 128;
    const int lanes = kSortedThreads / (slice / 8);
    const size_t smem = (size_t)3 * kSortedChunkBytes + (size_t)(lanes + 2) * (slice / 8) * NV * 4 + lanes * 4 +
                        ((size_t)R + 1 + P) * 4 + ((size_t)P / kRowsPerLane + 2) * 2;
    if (smem <= 200 * 1024) {
      static bool configured = false;  // per instantiation
      if (!configured) {
        FAVIT_CHECK_CUDA(cudaFuncSetAttribute(sppp_pool_fwd_sorted_kernel<TIn, TOut, 128>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        FAVIT_CHECK_CUDA(cudaFuncSetAttribute(sppp_pool_fwd_sorted_kernel<TIn, TOut, 64>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
      }
      int per_sm = (int)((220 * 1024) / (smem + 1024));
      per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
      const int sl_items = B * ceil_div(D * (int)sizeof(TIn), slice);
      const int grid = sl_items < num_sms() * per_sm ? sl_items : num_sms() * per_sm;
      note_kernel("sppp_pool_fwd_sorted_kernel<%d> grid=%d items=%d", slice, grid, sl_items);
      if (narrow)
        sppp_pool_fwd_sorted_kernel<TIn, TOut, 64><<<grid, kSortedThreads, smem, st>>>(
            (const TIn*)x, order, offsets, num_slots, (TOut*)out, P, R, D, r_cap, sl_items / B, sl_items);
      else
        sppp_pool_fwd_sorted_kernel<TIn, TOut, 128><<<grid, kSortedThreads, smem, st>>>(
            (const TIn*)x, order, offsets, num_slots, (TOut*)out, P, R, D, r_cap, sl_items / B, sl_items);
      FAVIT_CHECK_LAUNCH();
      return FAVIT_OK;
    }
  }
  const bool vec = (D % 8 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0);
  const int per = vec ? 256 : 32;
  const int64_t items = (int64_t)B * R * ceil_div(D, per);
  const unsigned blocks = (unsigned)ceil_div64(items, 8);
  note_kernel("sppp_pool_fwd_kernel<vec=%d>", vec ? 8 : 1);
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
  // (a) whole rows per CTA when the image's R gradient rows fit in shared memory (every model configuration)
  {
    constexpr int NVO = 16 / (int)sizeof(TOut);
    const size_t gbytes = (size_t)R * D * sizeof(TOut);
    const bool in_vec = sizeof(TIn) == 4 || (((uintptr_t)dout % 16 == 0) && (D % 8 == 0));
    if (((uintptr_t)dx % 16 == 0) && (D % NVO == 0) && in_vec && gbytes <= 96 * 1024 && D / NVO <= kBwdRowsThreads) {
      const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / (gbytes + 2048)));
      int nchunks = ceil_div(per_sm * num_sms(), B);                 // enough CTAs to fill the machine once ...
      nchunks = std::max(1, std::min(nchunks, ceil_div(P, 16)));     // ... with at least 16 rows each
      const int rows = ceil_div(P, nchunks);
      nchunks = ceil_div(P, rows);
      const size_t smem_r = gbytes + (size_t)rows * 4;
      if ((int64_t)B * nchunks < INT_MAX) {
        static bool configured_r = false;  // per instantiation
        if (!configured_r) {
          FAVIT_CHECK_CUDA(cudaFuncSetAttribute(sppp_pool_bwd_rows_kernel<TIn, TOut>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
          configured_r = true;
        }
        note_kernel("sppp_pool_bwd_rows_kernel grid=%d rows=%d", B * nchunks, rows);
        FAVIT_CHECK_CUDA(launch_pdl(sppp_pool_bwd_rows_kernel<TIn, TOut>, dim3((unsigned)(B * nchunks)), dim3(kBwdRowsThreads),
                                    smem_r, st, (const TIn*)dout, slot, counts, (TOut*)dx, P, R, D, r_cap, nchunks, rows));
        FAVIT_CHECK_LAUNCH();
        return FAVIT_OK;
      }
    }
  }
  const size_t smem = (size_t)R * kBwdCols * sizeof(float) + (size_t)P * 4;
  const int nslices = ceil_div(D, kBwdCols);
  if (((uintptr_t)dx % 16 == 0) && (((size_t)D * sizeof(TOut)) % 16 == 0) && smem <= 160 * 1024 &&
      (int64_t)B * nslices < INT_MAX) {
    static bool configured = false;  // per instantiation
    if (!configured) {
      FAVIT_CHECK_CUDA(cudaFuncSetAttribute(sppp_pool_bwd_tile_kernel<TIn, TOut>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      configured = true;
    }
    note_kernel("sppp_pool_bwd_tile_kernel grid=%d", B * nslices);
    sppp_pool_bwd_tile_kernel<TIn, TOut><<<(unsigned)(B * nslices), kBwdThreads, smem, st>>>(
        (const TIn*)dout, slot, counts, (TOut*)dx, P, R, D, r_cap, nslices);
    FAVIT_CHECK_LAUNCH();
    return FAVIT_OK;
  }
  const bool vec = (D % 8 == 0) && ((uintptr_t)dout % 16 == 0) && ((uintptr_t)dx % 16 == 0);
  const int per = vec ? 256 : 32;
  const int64_t items = (int64_t)B * P * ceil_div(D, per);
  const unsigned blocks = (unsigned)ceil_div64(items, 8);
  note_kernel("sppp_pool_bwd_kernel<vec=%d>", vec ? 8 : 1);
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

static int sppp_assign_impl(const int64_t* labels, int B, int img_h, int img_w, int patch, int grid,
                            int64_t* dom, int32_t* slot, int32_t* num_slots, int32_t* counts,
                            int64_t* slot_label, int32_t* offsets, int32_t* order, int r_cap, int K,
                            unsigned long long* acc, float* centroids, favit_stream stream) {
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
  const bool vec = (img_w % 2 == 0) && ((uintptr_t)labels % 16 == 0);
  // centroids in the same pass: the patches must tile the whole image, and the coordinate sums of a warp stay in int32
  const bool cent = acc && centroids && vec && (patch == 8 || patch == 16 || patch == 32) && img_h == grid * patch &&
                    img_w == grid * patch && img_w <= 32768 && img_h <= 32768 && K > 0 && K <= 4096;
  bool cent_done = false;
  if (cent) {
    FAVIT_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)B * 3 * K * sizeof(unsigned long long), st));
    if (patch == 16) sppp_dominant_vec_kernel<16, true><<<blocks, 256, 0, st>>>(labels, dom, B, img_h, img_w, grid, acc, K);
    else if (patch == 8) sppp_dominant_vec_kernel<8, true><<<blocks, 256, 0, st>>>(labels, dom, B, img_h, img_w, grid, acc, K);
    else sppp_dominant_vec_kernel<32, true><<<blocks, 256, 0, st>>>(labels, dom, B, img_h, img_w, grid, acc, K);
    cent_done = true;
    note_kernel("sppp_dominant_vec_kernel<%d, centroids=1>", patch);
  } else if (vec && patch == 16) {
    note_kernel("sppp_dominant_vec_kernel<16, centroids=0>");
    sppp_dominant_vec_kernel<16, false><<<blocks, 256, 0, st>>>(labels, dom, B, img_h, img_w, grid, nullptr, 0);
  } else if (vec && patch == 8) {
    note_kernel("sppp_dominant_vec_kernel<8, centroids=0>");
    sppp_dominant_vec_kernel<8, false><<<blocks, 256, 0, st>>>(labels, dom, B, img_h, img_w, grid, nullptr, 0);
  } else if (vec && patch == 32) {
    note_kernel("sppp_dominant_vec_kernel<32, centroids=0>");
    sppp_dominant_vec_kernel<32, false><<<blocks, 256, 0, st>>>(labels, dom, B, img_h, img_w, grid, nullptr, 0);
  } else {
    note_kernel("sppp_dominant_kernel (scalar loads, patch %d)", patch);
#define FAVIT_DOM(PPL)                                                                                   \
  sppp_dominant_kernel<PPL><<<blocks, 256, 0, st>>>(labels, dom, B, img_h, img_w, patch, grid)
    if (ppl <= 1) FAVIT_DOM(1);
    else if (ppl <= 2) FAVIT_DOM(2);
    else if (ppl <= 4) FAVIT_DOM(4);
    else if (ppl <= 8) FAVIT_DOM(8);
    else if (ppl <= 16) FAVIT_DOM(16);
    else FAVIT_DOM(32);
#undef FAVIT_DOM
  }
  FAVIT_CHECK_LAUNCH();
  const size_t slot_smem = (size_t)((P + 3) & ~3) * 8 + (size_t)P * 24 + 34 * 4;
  if (slot_smem <= 200 * 1024) {
    static bool configured = false;
    if (!configured) {
      FAVIT_CHECK_CUDA(cudaFuncSetAttribute(sppp_slot_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            200 * 1024));
      configured = true;
    }
    const int threads = P >= 1024 ? 1024 : ((P + 31) / 32) * 32;
    sppp_slot_smem_kernel<<<B, threads, slot_smem, st>>>(dom, slot, num_slots, counts, slot_label, offsets, order,
                                                        P, r_cap);
  } else if (P <= 1024) {
    sppp_slot_kernel<256><<<B, 256, 0, st>>>(dom, slot, num_slots, counts, slot_label, offsets, order, P, r_cap);
  } else {
    sppp_slot_kernel<1024><<<B, 1024, 0, st>>>(dom, slot, num_slots, counts, slot_label, offsets, order, P, r_cap);
  }
  FAVIT_CHECK_LAUNCH();
  if (acc && centroids) {
    if (!cent_done) return favit_sppp_centroids(labels, B, img_h, img_w, K, acc, centroids, stream);  // separate pass
    sppp_centroid_fin_kernel<<<(unsigned)ceil_div(B * K, 256), 256, 0, st>>>(acc, B, K, img_h, img_w, centroids);
    FAVIT_CHECK_LAUNCH();
  }
  return FAVIT_OK;
}

extern "C" int favit_sppp_assign(const int64_t* labels, int B, int img_h, int img_w, int patch, int grid,
                                 int64_t* dom, int32_t* slot, int32_t* num_slots, int32_t* counts,
                                 int64_t* slot_label, int32_t* offsets, int32_t* order, int r_cap,
                                 favit_stream stream) {
  return sppp_assign_impl(labels, B, img_h, img_w, patch, grid, dom, slot, num_slots, counts, slot_label, offsets, order,
                          r_cap, 0, nullptr, nullptr, stream);
}

extern "C" int favit_sppp_assign_centroids(const int64_t* labels, int B, int img_h, int img_w, int patch, int grid,
                                           int64_t* dom, int32_t* slot, int32_t* num_slots, int32_t* counts,
                                           int64_t* slot_label, int32_t* offsets, int32_t* order, int r_cap, int K,
                                           unsigned long long* acc, float* centroids, favit_stream stream) {
  FAVIT_CHECK_ARG(acc && centroids && K > 0, "sppp_assign_centroids: null accumulator / output or K <= 0");
  return sppp_assign_impl(labels, B, img_h, img_w, patch, grid, dom, slot, num_slots, counts, slot_label, offsets, order,
                          r_cap, K, acc, centroids, stream);
}

extern "C" int favit_sppp_centroids(const int64_t* labels, int B, int img_h, int img_w, int K, unsigned long long* acc,
                                    float* centroids, favit_stream stream) {
  FAVIT_CHECK_ARG(labels && acc && centroids, "sppp_centroids: null pointer");
  FAVIT_CHECK_ARG(B > 0 && img_h > 0 && img_w > 0 && K > 0, "sppp_centroids: sizes must be > 0");
  FAVIT_CHECK_ARG(K <= 4096, "sppp_centroids: K = %d > 4096 unsupported", K);
  FAVIT_CHECK_ARG((int64_t)img_h * img_w * (int64_t)(img_w + img_h) < ((int64_t)1 << 62), "sppp_centroids: image too large");
  cudaStream_t st = (cudaStream_t)stream;
  FAVIT_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)B * 3 * K * sizeof(unsigned long long), st));
  // per-CTA partial sums stay below 2^31: at most 32768 pixels (x, y < 65536 -> sums < 2^31) per CTA
  FAVIT_CHECK_ARG(img_w < 65536 && img_h < 65536, "sppp_centroids: image side must be < 65536");
  int rows = 32768 / img_w;
  rows = rows < 1 ? 1 : rows;
  const int want = ceil_div(8 * num_sms(), B);  // enough CTAs to fill the machine
  rows = std::min(rows, std::max(1, ceil_div(img_h, want)));
  const int chunks = ceil_div(img_h, rows);
  FAVIT_CHECK_ARG((int64_t)B * chunks < INT_MAX, "sppp_centroids: grid too large");
  sppp_centroid_acc_kernel<<<(unsigned)(B * chunks), 256, (size_t)3 * K * sizeof(unsigned), st>>>(labels, B, img_h, img_w, K,
                                                                                          rows, chunks, acc);
  FAVIT_CHECK_LAUNCH();
  sppp_centroid_fin_kernel<<<(unsigned)ceil_div(B * K, 256), 256, 0, st>>>(acc, B, K, img_h, img_w, centroids);
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

extern "C" int favit_sppp_pool_max_fwd(const void* x, favit_dtype x_dtype, const int32_t* order, const int32_t* offsets,
                                       const int32_t* num_slots, float* out, int32_t* argmax, int B, int P, int R, int D,
                                       int r_cap, favit_stream stream) {
  FAVIT_CHECK_ARG(x && order && offsets && num_slots && out && argmax, "sppp_pool_max_fwd: null pointer");
  FAVIT_CHECK_ARG(B > 0 && P > 0 && R > 0 && D > 0 && r_cap > 0 && (int64_t)B * R < INT_MAX, "sppp_pool_max_fwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == FAVIT_BF16)
    sppp_pool_max_fwd_kernel<__nv_bfloat16><<<(unsigned)(B * R), kVarThreads, 0, st>>>((const __nv_bfloat16*)x, order, offsets,
                                                                                    num_slots, out, argmax, P, R, D, r_cap);
  else if (x_dtype == FAVIT_F32)
    sppp_pool_max_fwd_kernel<float><<<(unsigned)(B * R), kVarThreads, 0, st>>>((const float*)x, order, offsets, num_slots, out,
                                                                            argmax, P, R, D, r_cap);
  else { set_error("sppp_pool_max_fwd: bad dtype"); return FAVIT_ERR_ARG; }
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

extern "C" int favit_sppp_pool_max_bwd(const float* dout, const int32_t* argmax, void* dx, favit_dtype dx_dtype, int B, int P,
                                       int R, int D, favit_stream stream) {
  FAVIT_CHECK_ARG(dout && argmax && dx, "sppp_pool_max_bwd: null pointer");
  FAVIT_CHECK_ARG(B > 0 && P > 0 && R > 0 && D > 0, "sppp_pool_max_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)B * R * D;
  const size_t es = dx_dtype == FAVIT_BF16 ? 2 : 4;
  FAVIT_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)B * P * D * es, st));
  const unsigned blocks = (unsigned)ceil_div64(total, 256);
  if (dx_dtype == FAVIT_BF16)
    sppp_pool_max_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(dout, argmax, (__nv_bfloat16*)dx, total, P, R, D);
  else if (dx_dtype == FAVIT_F32)
    sppp_pool_max_bwd_kernel<float><<<blocks, 256, 0, st>>>(dout, argmax, (float*)dx, total, P, R, D);
  else { set_error("sppp_pool_max_bwd: bad dtype"); return FAVIT_ERR_ARG; }
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

extern "C" int favit_sppp_pool_attn_fwd(const void* x, favit_dtype x_dtype, const int32_t* order, const int32_t* offsets,
                                        const int32_t* num_slots, float* out, float* weights, int B, int P, int R, int D,
                                        int r_cap, favit_stream stream) {
  FAVIT_CHECK_ARG(x && order && offsets && num_slots && out && weights, "sppp_pool_attn_fwd: null pointer");
  FAVIT_CHECK_ARG(B > 0 && P > 0 && R > 0 && D > 0 && r_cap > 0 && (int64_t)B * R < INT_MAX, "sppp_pool_attn_fwd: bad sizes");
  FAVIT_CHECK_ARG(P <= 12000, "sppp_pool_attn_fwd: more than 12000 patches per image unsupported");
  cudaStream_t st = (cudaStream_t)stream;
  FAVIT_CHECK_CUDA(cudaMemsetAsync(weights, 0, (size_t)B * P * sizeof(float), st));  // patches outside every kept slot
  const size_t smem = (size_t)P * sizeof(float);
  if (x_dtype == FAVIT_BF16)
    sppp_pool_attn_fwd_kernel<__nv_bfloat16><<<(unsigned)(B * R), kVarThreads, smem, st>>>(
        (const __nv_bfloat16*)x, order, offsets, num_slots, out, weights, P, R, D, r_cap);
  else if (x_dtype == FAVIT_F32)
    sppp_pool_attn_fwd_kernel<float><<<(unsigned)(B * R), kVarThreads, smem, st>>>((const float*)x, order, offsets, num_slots,
                                                                                out, weights, P, R, D, r_cap);
  else { set_error("sppp_pool_attn_fwd: bad dtype"); return FAVIT_ERR_ARG; }
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

extern "C" int favit_sppp_pool_attn_bwd(const void* x, favit_dtype x_dtype, const float* dout, const float* out,
                                        const float* weights, const int32_t* order, const int32_t* offsets,
                                        const int32_t* num_slots, void* dx, int B, int P, int R, int D, int r_cap,
                                        favit_stream stream) {
  FAVIT_CHECK_ARG(x && dout && out && weights && order && offsets && num_slots && dx, "sppp_pool_attn_bwd: null pointer");
  FAVIT_CHECK_ARG(B > 0 && P > 0 && R > 0 && D > 0 && r_cap > 0 && (int64_t)B * R < INT_MAX, "sppp_pool_attn_bwd: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t es = x_dtype == FAVIT_BF16 ? 2 : 4;
  FAVIT_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)B * P * D * es, st));
  if (x_dtype == FAVIT_BF16)
    sppp_pool_attn_bwd_kernel<__nv_bfloat16><<<(unsigned)(B * R), kVarThreads, 0, st>>>(
        (const __nv_bfloat16*)x, dout, out, weights, order, offsets, num_slots, (__nv_bfloat16*)dx, P, R, D, r_cap);
  else if (x_dtype == FAVIT_F32)
    sppp_pool_attn_bwd_kernel<float><<<(unsigned)(B * R), kVarThreads, 0, st>>>((const float*)x, dout, out, weights, order,
                                                                             offsets, num_slots, (float*)dx, P, R, D, r_cap);
  else { set_error("sppp_pool_attn_bwd: bad dtype"); return FAVIT_ERR_ARG; }
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
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
