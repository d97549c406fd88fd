// Device and host helpers shared by the whole-sequence (mhla_window_attn_seq.cu) and chunked (mhla_window_attn_chunk.cu)
// TMA + mma.sync window-attention kernels: swizzled-tile addressing, the row-mapped score / PV / dX tile products, the
// aligned key-slot numbering of the backward pass, column sums on the tensor pipe, and the 4-D tensor map over the
// packed [B, N, 3, H, 64] qkv tensor.
#pragma once
#include <cuda.h>

#include <mutex>

#include "attn_mma_common.cuh"
#include "tcgen05_ptx.cuh"

namespace favit {
namespace seqk {

using namespace attn;

constexpr int HD = 64;
constexpr int kRowBytes = HD * 2;
constexpr int kMaxThreads = 832;  // 26 tiles; keeps 78 registers per thread available at two CTAs of 13 warps per SM

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(tm), "r"(c0), "r"(c1), "r"(c2),
               "r"(c3)
               : "memory");
}

// byte offset of 16-byte chunk `chunk` of row `row` in a 128-byte-row tile with the 128B swizzle (base 1024-aligned)
__device__ __forceinline__ uint32_t row_off(int row, int chunk) { return (uint32_t)((row * 8 + (chunk ^ (row & 7))) * 16); }

// acc[NT][4] = A(own 16-row tile) . B^T where B row of slot s is sequence row brow[.] (per-lane, see load_b)
template <int NT>
__device__ __forceinline__ void scores_rows(const uint8_t* sA_tile, const uint8_t* sB, const int (&brow)[NT / 2], int lane,
                                            float (&acc)[NT][4]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
  const uint32_t bbase = smem_u32(sB);
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    uint32_t a[4];
    load_a<HD>(sA_tile, ks, lane, a);
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b[4];
      ldsm_x4(bbase + row_off(brow[np], 2 * ks + ((lane >> 3) & 1)), b);
      mma_bf16(acc[2 * np], a, b[0], b[1]);
      mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
    }
  }
}
// acc[8][4] += P(16 x 8NT, accumulator fragments) . rows, where the row of slot s is sequence row rrow[.] (see load_bt)
template <int NT>
__device__ __forceinline__ void pv_rows(const float (&p)[NT][4], const uint8_t* sRows, const int (&rrow)[NT / 2], int lane,
                                        float (&acc)[HD / 8][4]) {
  const uint32_t rbase = smem_u32(sRows);
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
      ldsm_x4_trans(rbase + row_off(rrow[kk], nd + (lane >> 4)), b);
      mma_bf16(acc[nd], a, b[0], b[1]);
      mma_bf16(acc[nd + 1], a, b[2], b[3]);
    }
  }
}

// Phase A numbers the key slots of query tile i0 in ALIGNED coordinates: slot s <-> key i0 - 8 + s for s in 1..30 (the
// band i0-h .. i0+15+h fits for h <= 7), slot 0 <-> the duplicated edge key N-1, slot 31 <-> the duplicated edge key 0.
// A duplicate goes to its edge slot unless the edge key lies inside the query's OWN window (then the band slot carries
// multiplicity 1 + pad), so every non-edge entry of P / dS sits within h of the diagonal.  The 16 x 32 blocks P and dS
// of every tile are then 8-aligned in key space: a key tile finds the six 8 x 8 blocks that touch it at fixed chunk
// positions of its own and its two neighbour query tiles, transposes them with ldmatrix.trans, and never recomputes a
// score; the first / last key tile add one k-step for what the edge slots hold.
struct AlSlots {
  int base, lo, hi, exA, exB, N;
  __device__ __forceinline__ int key(int s) const {
    if (s == 0) return exA ? N - 1 : -1;
    if (s == 31) return exB ? 0 : -1;
    const int j = base + s;
    return (j >= lo && j <= hi) ? j : -1;
  }
};
__device__ __forceinline__ AlSlots al_slots(int i0, int N, int W) {
  const int h = W >> 1;
  AlSlots k;
  k.N = N;
  k.base = i0 - 8;
  k.lo = max(0, i0 - h);
  k.hi = min(N - 1, i0 + 15 + h);
  k.exA = (i0 == 0) ? 1 : 0;  // early queries (i <= h, the only ones that duplicate key N-1) live in the first tile
  k.exB = 1;
  return k;
}
__device__ __forceinline__ RowSlots al_row_slots(const AlSlots& k, int i, int N, int W) {
  const WindowRow r = window_row(i, N, W);
  RowSlots o;
  o.s_lo = r.s - k.base;
  o.s_span = r.e - 1 - r.s;
  if (r.tgt >= r.s && r.tgt < r.e) o.ts = r.tgt - k.base;       // inside the query's own window: band slot
  else o.ts = (r.s == 0) ? (k.exA ? 0 : -1) : 31;               // else the edge slot of key N-1 / key 0
  if (r.pad == 0) o.ts = -1;
  o.lb_in = kLog2Int[1 + r.pad];
  o.lb_out = kLog2Int[r.pad];
  return o;
}

constexpr int kBlkBytes = 16 * 32 * 2;  // one parked 16 x 32 bf16 block (rows of 64 bytes, tile_off<32> swizzle)

// acc[8][4] += A . rows(16 x 64), A given as ready fragments, B rows = sequence rows brow (load_bt lane pattern)
__device__ __forceinline__ void mma_rows(const uint32_t (&a)[4], const uint8_t* sRows, int brow, int lane,
                                         float (&acc)[HD / 8][4]) {
  const uint32_t rbase = smem_u32(sRows);
#pragma unroll
  for (int nd = 0; nd < HD / 8; nd += 2) {
    uint32_t b[4];
    ldsm_x4_trans(rbase + row_off(brow, nd + (lane >> 4)), b);
    mma_bf16(acc[nd], a, b[0], b[1]);
    mma_bf16(acc[nd + 1], a, b[2], b[3]);
  }
}

// Column-sum accumulation, two adjacent columns at a time.  shared: the CTA's accumulator (fp32 shared-memory adds are
// CAS loops in SASS: cheap only while few warps meet); global: one 8-byte vector reduction per pair.
__device__ __forceinline__ void add_pair(float* dst, float a, float b, bool shared) {
  if (shared) {
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(smem_u32(dst)), "f"(a) : "memory");
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(smem_u32(dst + 1)), "f"(b) : "memory");
  } else {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(a), "f"(b) : "memory");
  }
}

// Column sums on the tensor pipe: ones(16 x 16) . X(16 x 64) leaves the 64 sums in every row of the accumulator; lanes
// 0-3 (row 0) add them to dst (shared memory).  X = a staged tile (bf16 rows in shared memory) ...
__device__ __forceinline__ void colsum_staged(const uint8_t* stage, int lane, float* dst, bool shared) {
  const uint32_t ones[4] = {kOnesBf16x2, kOnesBf16x2, kOnesBf16x2, kOnesBf16x2};
  float acc[HD / 8][4];
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
  mma_rows(ones, stage, (lane & 7) + 8 * ((lane >> 3) & 1), lane, acc);
  if (lane < 4) {
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd) add_pair(dst + nd * 8 + lane * 2, acc[nd][0], acc[nd][1], shared);
  }
}
// ... or X = packed bf16 accumulator fragments v0 (rows 0-7) / v1 (rows 8-15) of each 8-column step: movmatrix turns
// them into the B operand (k = row, n = column) without leaving the registers.
__device__ __forceinline__ void colsum_frags(const uint32_t (&v0)[HD / 8], const uint32_t (&v1)[HD / 8], int lane, float* dst,
                                             bool shared) {
  const uint32_t ones[4] = {kOnesBf16x2, kOnesBf16x2, kOnesBf16x2, kOnesBf16x2};
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    mma_bf16(acc, ones, movmatrix_trans(v0[nd]), movmatrix_trans(v1[nd]));
    if (lane < 4) add_pair(dst + nd * 8 + lane * 2, acc[0], acc[1], shared);
  }
}

// ---- host side: tensor maps ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
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

// [B][N][H][64] bf16 view with element strides (sb, sn, sh); box = 64 x 1 x rows_box x 1
inline int make_map(CUtensorMap* tm, const void* ptr, int B, int H, int N, int64_t sb, int64_t sn, int64_t shh, int rows_box) {
  cuuint64_t gdim[4] = {(cuuint64_t)HD, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)shh * 2, (cuuint64_t)sn * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {(cuuint32_t)HD, 1, (cuuint32_t)rows_box, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attn_seq: cuTensorMapEncodeTiled failed with CUresult %d (B=%d H=%d N=%d sb=%lld sn=%lld sh=%lld)", (int)r,
              B, H, N, (long long)sb, (long long)sn, (long long)shh);
    return FAVIT_ERR_CUDA;
  }
  return FAVIT_OK;
}

}  // namespace seqk
}  // namespace favit
