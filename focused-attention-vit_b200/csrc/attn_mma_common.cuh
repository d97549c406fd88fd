// Device helpers shared by the tensor-core-tile window-attention kernels (mhla_window_attn_mma.cu: one warp stages its
// own tile; mhla_window_attn_chunk.cu: a CTA stages a whole chunk of the sequence once).
#pragma once
#include <math_constants.h>

#include "favit_common.cuh"

namespace favit {
namespace attn {


// log2 of the window multiplicities 0..16 (windows up to 15 on the tensor-core paths); [0] = -inf
__constant__ float kLog2Int[17] = {-__builtin_huge_valf(), 0.f, 1.f, 1.5849625007211562f, 2.f, 2.321928094887362f,
                                   2.584962500721156f, 2.807354922057604f, 3.f, 3.169925001442312f,
                                   3.321928094887362f, 3.4594316186372978f, 3.584962500721156f, 3.700439718141092f,
                                   3.807354922057604f, 3.9068905956085187f, 4.f};

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct Shape {
  int B, H, N, W;
  int64_t sb, sn, sh;  // q/k/v strides in elements (shared by dq/dk/dv)
  float scale_log2, scale;
  int tiles;           // 16-row tiles per sequence
  float* colsum;       // backward only, may be null: [3 * H * hd] fp32 column sums of dq | dk | dv (accumulated)
};

struct WindowRow {
  int s, e, pad, tgt;
};
__device__ __forceinline__ WindowRow window_row(int i, int N, int W) {
  const int h = W >> 1;
  WindowRow r;
  r.s = max(0, i - h);
  r.e = min(N, i + h + 1);
  r.pad = max(0, W - (r.e - r.s));
  r.tgt = (r.s == 0) ? (N - 1) : 0;
  return r;
}
__device__ __forceinline__ int window_mult(const WindowRow& r, int j) {
  return ((j >= r.s && j < r.e) ? 1 : 0) + ((j == r.tgt) ? r.pad : 0);
}

// ---- shared-memory tile of rows of HD bf16, 16-byte chunks XOR-swizzled so that ldmatrix is conflict-free --------
template <int HD>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  constexpr int CH = HD / 8;
  const int swz = (CH >= 8) ? (row & 7) : ((row >> 1) & (CH - 1));
  return (uint32_t)((row * CH + (chunk ^ swz)) * 16);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // 0 -> the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ float bf16x2_dot(uint32_t a, uint32_t b) {
  return __uint_as_float(a << 16) * __uint_as_float(b << 16) +
         __uint_as_float(a & 0xffff0000u) * __uint_as_float(b & 0xffff0000u);
}

// Stage `rows` rows of HD bf16 into a swizzled tile: row r comes from src_row(r) (nullptr -> zero fill).
template <int HD, typename F>
__device__ __forceinline__ void stage_rows(uint8_t* tile, int rows, int lane, const __nv_bfloat16* safe, F src_row) {
  constexpr int CH = HD / 8;
  const uint32_t base = smem_u32(tile);
  for (int idx = lane; idx < rows * CH; idx += 32) {
    const int r = idx / CH, c = idx % CH;
    const __nv_bfloat16* src = src_row(r);
    // a zero-size copy still carries an address: keep it a valid global one
    cp_async16(base + tile_off<HD>(r, c), src ? (const void*)(src + c * 8) : (const void*)safe, src != nullptr);
  }
}

// A fragments (16 rows x 16 k) of k-step ks from a 16-row tile.
template <int HD>
__device__ __forceinline__ void load_a(const uint8_t* tile, int ks, int lane, uint32_t (&a)[4]) {
  const int row = (lane & 7) + 8 * ((lane >> 3) & 1);
  ldsm_x4(smem_u32(tile) + tile_off<HD>(row, 2 * ks + (lane >> 4)), a);
}
// B fragments for two adjacent n-tiles (rows 8nt.. and 8nt+8..) of k-step ks; smem rows index n, k is contiguous.
template <int HD>
__device__ __forceinline__ void load_b(const uint8_t* tile, int nt, int ks, int lane, uint32_t (&b)[4]) {
  const int row = 8 * nt + (lane & 7) + 8 * (lane >> 4);
  ldsm_x4(smem_u32(tile) + tile_off<HD>(row, 2 * ks + ((lane >> 3) & 1)), b);
}
// B fragments for k-step kk (smem rows 16kk.. index k) and two adjacent n-tiles nd, nd+1 (8-column chunks).
template <int HD>
__device__ __forceinline__ void load_bt(const uint8_t* tile, int kk, int nd, int lane, uint32_t (&b)[4]) {
  const int row = 16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1);
  ldsm_x4_trans(smem_u32(tile) + tile_off<HD>(row, nd + (lane >> 4)), b);
}

// accumulators [HD/8][4] -> bf16 rows in a 16-row staging tile (thread holds rows lane/4 and lane/4 + 8)
template <int HD>
__device__ __forceinline__ void stage_acc(uint8_t* tile, const float (&acc)[HD / 8][4], float m0, float m1, int lane) {
  const int r0 = lane >> 2, r1 = r0 + 8, sub = (lane & 3) * 4;
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd) {
    *reinterpret_cast<uint32_t*>(tile + tile_off<HD>(r0, nd) + sub) = pack_bf16x2(acc[nd][0] * m0, acc[nd][1] * m0);
    *reinterpret_cast<uint32_t*>(tile + tile_off<HD>(r1, nd) + sub) = pack_bf16x2(acc[nd][2] * m1, acc[nd][3] * m1);
  }
}
// staged 16-row tile -> global rows (16-byte row segments); dst_row(r) == nullptr skips the row
template <int HD, typename F>
__device__ __forceinline__ void store_rows(const uint8_t* tile, int lane, F dst_row) {
  constexpr int CH = HD / 8;
#pragma unroll
  for (int idx = lane; idx < 16 * CH; idx += 32) {
    const int r = idx / CH, c = idx % CH;
    __nv_bfloat16* dst = dst_row(r);
    if (dst) *reinterpret_cast<uint4*>(dst + c * 8) = *reinterpret_cast<const uint4*>(tile + tile_off<HD>(r, c));
  }
}

// 8 x 8 b16 transpose across the warp: fragment (row lane/4, cols 2*(lane%4)+{0,1}) -> the same of the transposed matrix
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
constexpr uint32_t kOnesBf16x2 = 0x3f803f80u;  // {1.0, 1.0} in bf16

// column sums of a staged 16-row tile (rows past the end of the sequence hold exact zeros) added to dst[0..HD)
template <int HD>
__device__ __forceinline__ void tile_colsum(const uint8_t* tile, int lane, float* dst) {
  constexpr int CPL = HD / 32;  // columns per lane
  if (CPL == 0) return;
  float s[CPL > 0 ? CPL : 1];
#pragma unroll
  for (int c = 0; c < CPL; ++c) s[c] = 0.f;
  const int c0 = lane * CPL;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const uint8_t* p = tile + tile_off<HD>(r, c0 >> 3) + (c0 & 7) * 2;
#pragma unroll
    for (int c = 0; c < CPL; ++c) s[c] += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(p + 2 * c));
  }
#pragma unroll
  for (int c = 0; c < CPL; ++c) atomicAdd(dst + c0 + c, s[c]);
}

// ---- key slots of a query tile -----------------------------------------------------------------------------------
struct KeySlots {
  int lo, nband, exA, exB, N;
  __device__ __forceinline__ int key(int s) const {
    if (s < nband) return lo + s;
    if (s == nband) return exA ? N - 1 : -1;
    if (s == nband + 1) return exB ? 0 : -1;
    return -1;
  }
  __device__ __forceinline__ int mult(const WindowRow& r, int s) const {
    const int j = key(s);
    if (j < 0) return 0;
    if (s < nband) return window_mult(r, j);
    return (j == r.tgt) ? r.pad : 0;  // an edge row outside the band: only the duplicated index reaches it
  }
};
// Per-query view of a key-slot list: slots [s_lo, s_hi] are the band (multiplicity 1), slot `ts` is the duplicated edge
// key; log2 of the multiplicities is precomputed so that the per-element work is two compares and a select.
struct RowSlots {
  int s_lo, s_span, ts;
  float lb_in, lb_out;   // log2(1 + pad) when the edge key lies inside the band, log2(pad) (or -inf) when outside
  __device__ __forceinline__ float bias(int slot) const {
    const bool in_band = (unsigned)(slot - s_lo) <= (unsigned)s_span;
    const bool edge = slot == ts;
    return in_band ? (edge ? lb_in : 0.f) : (edge ? lb_out : -CUDART_INF_F);
  }
};
__device__ __forceinline__ RowSlots row_slots(const KeySlots& k, int i, int N, int W) {
  const WindowRow r = window_row(i, N, W);
  RowSlots o;
  o.s_lo = r.s - k.lo;
  o.s_span = r.e - 1 - r.s;
  const int hi = k.lo + k.nband - 1;
  if (r.tgt >= k.lo && r.tgt <= hi) o.ts = r.tgt - k.lo;
  else o.ts = (r.tgt == N - 1) ? (k.exA ? k.nband : -1) : (k.exB ? k.nband + 1 : -1);
  if (r.pad == 0) o.ts = -1;
  o.lb_in = kLog2Int[1 + r.pad];
  o.lb_out = kLog2Int[r.pad];
  return o;
}

__device__ __forceinline__ KeySlots key_slots(int i0, int N, int W) {
  const int h = W >> 1;
  KeySlots k;
  k.N = N;
  k.lo = max(0, i0 - h);
  const int hi = min(N - 1, i0 + 15 + h);
  k.nband = hi - k.lo + 1;
  k.exA = (N - 1 > hi) ? 1 : 0;
  k.exB = (k.lo > 0) ? 1 : 0;
  return k;
}

template <int HD, int NT>
__device__ __forceinline__ void scores(const uint8_t* sA, const uint8_t* sB, int lane, float (&acc)[NT][4]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    uint32_t a[4];
    load_a<HD>(sA, ks, lane, a);
#pragma unroll
    for (int nt = 0; nt < NT; nt += 2) {
      uint32_t b[4];
      load_b<HD>(sB, nt, ks, lane, b);
      mma_bf16(acc[nt], a, b[0], b[1]);
      mma_bf16(acc[nt + 1], a, b[2], b[3]);
    }
  }
}

// acc[HD/8][4] += P(16 x 8NT, from accumulator fragments) . rows(8NT x HD)
template <int HD, int NT>
__device__ __forceinline__ void pv(const float (&p)[NT][4], const uint8_t* sRows, int lane, float (&acc)[HD / 8][4]) {
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int nd = 0; nd < HD / 8; nd += 2) {
      uint32_t b[4];
      load_bt<HD>(sRows, kk, nd, lane, b);
      mma_bf16(acc[nd], a, b[0], b[1]);
      mma_bf16(acc[nd + 1], a, b[2], b[3]);
    }
  }
}


// ---- query slots of a key tile (dK / dV pass) ------------------------------------------------------------------
struct QuerySlots {
  int lo, nband, nA, b0, nB;  // band queries lo.., then nA early queries 0.. (reach key N-1), then nB late queries b0..
  __device__ __forceinline__ int query(int s) const {
    if (s < nband) return lo + s;
    s -= nband;
    if (s < nA) return s;
    s -= nA;
    if (s < nB) return b0 + s;
    return -1;
  }
};
__device__ __forceinline__ QuerySlots query_slots(int j0, int N, int W) {
  const int h = W >> 1;
  QuerySlots qs;
  qs.lo = max(0, j0 - h);
  const int hi = min(N - 1, j0 + 15 + h);
  qs.nband = hi - qs.lo + 1;
  // key N-1 in this tile: queries with s == 0 (i <= h) duplicate it when their window is short
  qs.nA = (j0 + 15 >= N - 1) ? min(qs.lo, h + 1) : 0;
  // key 0 in this tile: queries with s > 0 (i > h) whose window runs past the end (i >= N - h) duplicate it
  qs.b0 = max(max(hi + 1, N - h), h + 1);
  qs.nB = (j0 == 0) ? max(0, N - qs.b0) : 0;
  return qs;
}


}  // namespace attn
}  // namespace favit
