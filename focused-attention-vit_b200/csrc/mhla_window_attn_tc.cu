// MHLA window attention forward for WIDE windows (17 <= W <= 65) on the 5th-generation tensor cores:
// tcgen05.mma with TMEM accumulators, operands staged by TMA, the probabilities handed from the softmax back to the
// tensor core THROUGH TENSOR MEMORY (they never touch shared memory).  Replaces /root/reference/models/mhla.py:109-154
// (window index table, K/V window gathers, scaled scores, softmax, PV) for the window sizes where a 128-row query tile
// is dense enough to feed a 128 x KT MMA: at W = 63 a query tile needs KT = 192 keys and a third of the score block is
// inside the band (at the reference's default W = 7 it would be 5 %, which is why that case runs on 16 x 8 mma.sync
// tiles in mhla_window_attn_seq.cu).  head_dim 64, bf16, N >= W, no mask / dropout.
//
// One CTA = one (image, head, 128-query tile); 128 threads; thread r owns query row r = TMEM lane r.
//   TMA:    Q [128 x 64], K and V [KT x 64] (KT = 128 + 2h rounded up to 16; rows before key 0 / after key N-1 and the
//           rows of other sequences are zero-filled by the 4-D tensor map) -> shared memory, SWIZZLE_128B
//   MMA 1:  S[128 x KT] = Q . K^T          (kind::f16, A and B from shared memory, fp32 accumulator in TMEM)
//   softmax in registers: warp w reads only the 32 + 2h columns its 32 rows can reach (tcgen05.ld 32x32b), applies the band /
//           sequence-bound mask, adds the duplicated-edge term of mhla.py:71-79 (key N-1 with multiplicity pad for rows
//           clipped on the left, key 0 for rows clipped on the right: one extra 64-wide dot product on the CUDA cores for
//           the <= 2h affected rows of a sequence), exponentiates (exp2, scale folded in)
//   P -> TMEM: packed bf16 pairs over the first KT/2 columns of S (tcgen05.st; every lane has read its S row by then)
//   MMA 2:  O[128 x 64] = P . V            (A from TENSOR MEMORY, B = V from shared memory, MN-major)
//   epilogue: O / row sum (+ edge term) -> bf16, 128 contiguous bytes per row; LSE for the backward pass.
// TMEM: KT + 64 <= 256 columns per CTA, two CTAs per SM.  The op stays HBM-bound (AI = W / 2 FLOP per byte); the tensor
// core is what keeps the contraction off the critical path that the CUDA-core kernel (4 N W D FMAs) sits on.
#include <math_constants.h>

#include <algorithm>
#include <mutex>

#include "favit_common.cuh"
#include "tcgen05_ptx.cuh"

namespace favit {
namespace {

using namespace ptx;

constexpr int HD = 64;
constexpr int BQ = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct TcParams {
  int B, H, N, W, h, KT, qtiles;
  float scale_log2;
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
  int64_t sb, sn, sh;   // element strides of q / k / v
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc1(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc1(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit1(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: A = 128 lanes x (K / 2) columns of packed bf16 pairs
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// The softmax of one warp (rows 32 w .. 32 w + 31).  NCH = 32-column chunks of S the warp has to read: its rows reach
// columns 32 w .. 32 w + 31 + 2h.  Register k of a lane always holds column 32 w + k, so the code is the same for the four
// warps (only TMEM addresses depend on w): the first version specialised the register indices per warp and ran four
// 48 KB instruction streams per CTA — 'no instruction' was its dominant stall.
// Writes the unnormalised probabilities as bf16 pairs: this warp's 16 NCH packed columns at [16 w, 16 w + 16 NCH), zeros
// over the rest of the NPC = 96 packed columns the second MMA reads; returns the row maximum (log2 domain), the row sum
// and the edge probability of this thread's row.
template <int NCH>
__device__ __forceinline__ void softmax_rows(uint32_t tmem_base, const TcParams& p, int q0, int w, int lane, float t_edge,
                                             float& mx_out, float& sum_out, float& pe_out) {
  float v[NCH * 32];
  const uint32_t lane_base = tmem_base + ((uint32_t)(w * 32) << 16);
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t r[32];
    tmem_ld32(lane_base + (uint32_t)(w * 32 + c * 32), r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[c * 32 + i] = __uint_as_float(r[i]);
  }
  const int qi = q0 + w * 32 + lane;
  // register k <-> key qi - h + (k - lane): valid iff 0 <= k - lane <= 2h and the key lies in [0, N)
  const int lo = lane + max(0, p.h - qi);
  const int hi = lane + min(2 * p.h, p.N - 1 - qi + p.h);
  float mx = t_edge;                    // -inf when this row has no duplicated edge key
#pragma unroll
  for (int k = 0; k < NCH * 32; ++k) {
    const bool ok = k >= lo && k <= hi;
    v[k] = ok ? v[k] * p.scale_log2 : -CUDART_INF_F;
    mx = fmaxf(mx, v[k]);
  }
  if (qi >= p.N || hi < lo) mx = 0.f;   // rows past the sequence: all masked; keep the arithmetic finite
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NCH * 32; ++k) {
    v[k] = ex2f(v[k] - mx);
    sum += v[k];
  }
  const float pe = ex2f(t_edge - mx);
  sum += pe;
  // own columns: registers (2 i, 2 i + 1) -> packed column 16 w + i
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = pack_bf16x2(v[c * 32 + 2 * i], v[c * 32 + 2 * i + 1]);
    tmem_st16(lane_base + (uint32_t)(16 * w + 16 * c), r);
  }
  // the rest of the 96 packed columns: zeros (6 - NCH chunks of 16, before and after the warp's own range)
  {
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
    for (int t = 0; t < 6 - NCH; ++t) tmem_st16(lane_base + (uint32_t)(t < w ? 16 * t : 16 * (t + NCH)), z);
  }
  tmem_wait_st();
  mx_out = mx;
  sum_out = sum;
  pe_out = pe;
}

template <int NCH, int NPCH>
__global__ void __launch_bounds__(BQ, 2) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmq,
                                                            const __grid_constant__ CUtensorMap tmk,
                                                            const __grid_constant__ CUtensorMap tmv,
                                                            __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                                                            const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t sQ = base, sK = sQ + BQ * 128, sV = sK + (uint32_t)p.KT * 128;
  const uint32_t misc = sV + (uint32_t)p.KT * 128;                  // barriers, TMEM slot, edge rows
  const uint32_t bar_load = misc, bar_s = misc + 8, bar_o = misc + 16, tmem_slot = misc + 24;
  uint8_t* g_misc = gen + (misc - base);
  // edge rows: [0] k_0, [1] k_{N-1}, [2] v_0, [3] v_{N-1}, 128 bytes each
  __nv_bfloat16* s_edge = reinterpret_cast<__nv_bfloat16*>(g_misc + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x % p.qtiles;
  const int bh = blockIdx.x / p.qtiles;
  const int hh = bh % p.H, b = bh / p.H;
  const int q0 = tile * BQ, kb = q0 - p.h;

  if (tid == 0) {
    prefetch_tmap(&tmq);
    prefetch_tmap(&tmk);
    prefetch_tmap(&tmv);
    mbar_init(bar_load, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
    mbar_expect_tx(bar_load, (uint32_t)(BQ + 2 * p.KT) * 128u);
    tma_load_4d(sQ, &tmq, bar_load, 0, hh, q0, b);
    tma_load_4d(sK, &tmk, bar_load, 0, hh, kb, b);
    tma_load_4d(sV, &tmv, bar_load, 0, hh, kb, b);
  }
  if (warp == 1) tmem_alloc1(tmem_slot, 256);
  if (warp >= 2) {  // the four edge rows: 32 x 16 bytes, two warps x 16 lanes
    const int t = tid - 64;
    if (t < 32) {
      const int which = t >> 3, chunk = t & 7;
      const __nv_bfloat16* src = (which < 2 ? p.k : p.v) + (int64_t)b * p.sb + (int64_t)hh * p.sh +
                                 (int64_t)((which & 1) ? p.N - 1 : 0) * p.sn + chunk * 8;
      *reinterpret_cast<uint4*>(s_edge + which * 64 + chunk * 8) = *reinterpret_cast<const uint4*>(src);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(g_misc + 24);
  const uint32_t tmem_o = tmem_base + (uint32_t)p.KT;

  if (tid == 0) {
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(BQ, p.KT, 0, 0);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k)
      umma_ss(tmem_base, make_smem_desc(sQ + k * 32u, 16u, 1024u), make_smem_desc(sK + k * 32u, 16u, 1024u), idesc,
              k > 0 ? 1u : 0u);
    umma_commit1(bar_s);
  }

  // ---- this row's duplicated-edge logit (mhla.py:71-79), while the tensor core works ----
  const int qi = q0 + tid;
  float t_edge = -CUDART_INF_F;
  int edge_row = -1;   // 0: key 0, 1: key N-1
  {
    const int s = max(0, qi - p.h), e = min(p.N, qi + p.h + 1);
    const int pad = p.W - (e - s);
    if (qi < p.N && pad > 0) {
      edge_row = (s == 0) ? 1 : 0;
      mbar_wait(bar_load, 0);   // Q is in shared memory (generic-proxy reads of async-proxy writes: ordered by the barrier)
      const __nv_bfloat16* ke = s_edge + edge_row * 64;
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 qv = *reinterpret_cast<const uint4*>(gen + (sQ - base) + tid * 128 + ((c ^ (tid & 7)) << 4));
        const uint4 kv = *reinterpret_cast<const uint4*>(ke + c * 8);
        const uint32_t qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          dot = fmaf(__uint_as_float(qa[t] << 16), __uint_as_float(ka[t] << 16), dot);
          dot = fmaf(__uint_as_float(qa[t] & 0xffff0000u), __uint_as_float(ka[t] & 0xffff0000u), dot);
        }
      }
      t_edge = dot * p.scale_log2 + log2f((float)pad);
    }
  }

  mbar_wait(bar_s, 0);
  tc_fence_after();
  float mx, sum, pe;
  softmax_rows<NCH>(tmem_base, p, q0, warp, lane, t_edge, mx, sum, pe);
  tc_fence_before();
  __syncthreads();   // every lane's P row is in tensor memory
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(BQ, HD, 0, 1);   // B = V: [keys][64], the N dimension contiguous
    for (int kk = 0; kk < p.KT / 16; ++kk)
      umma_ts(tmem_o, tmem_base + (uint32_t)(kk * 8), make_smem_desc(sV + kk * 2048u, 8192u, 1024u), idesc,
              kk > 0 ? 1u : 0u);
    umma_commit1(bar_o);
  }
  mbar_wait(bar_o, 0);
  tc_fence_after();
  {
    const bool live = qi < p.N;
    const float inv = live ? 1.f / sum : 0.f;
    const uint32_t lane_base = tmem_o + ((uint32_t)(warp * 32) << 16);
    __nv_bfloat16* orow = out + ((int64_t)(b * p.N + (live ? qi : 0)) * p.H + hh) * HD;
    const __nv_bfloat16* ve = s_edge + (2 + (edge_row < 0 ? 0 : edge_row)) * 64;
    const float pw = edge_row < 0 ? 0.f : pe;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(lane_base + (uint32_t)(c * 32), r);   // warp-collective: rows past the sequence take part too
      tmem_wait_ld();
      if (live) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            f[e] = (__uint_as_float(r[g8 * 8 + e]) + pw * __bfloat162float(ve[c * 32 + g8 * 8 + e])) * inv;
          store8(orow + c * 32 + g8 * 8, f);
        }
      }
    }
    if (live) lse[((int64_t)b * p.H + hh) * p.N + qi] = (mx + log2f(sum)) * kLn2;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc1(tmem_base, 256);
  }
}

// =====================================================================================================================
// Backward for the same window range, three kernels (autograd of mhla.py:109-154; SURVEY.md §8a closed form):
//   P_ij = m_ij exp(s_ij - L_i),  delta_i = dO_i . O_i,  dP_ij = dO_i . V_j,  dS_ij = P_ij (dP_ij - delta_i)
//   dQ_i = scale sum_j dS_ij K_j,   dK_j = scale sum_i dS_ij Q_i,   dV_j = sum_i P_ij dO_i
// (A) attn_tc_bwd_dq_kernel — query-major, one CTA per 128-query tile: S = Q.K^T and dP = dO.V^T on the tensor core (the
//     second overwrites the first in tensor memory once every lane has turned its S row into P), dS goes back to tensor
//     memory as packed bf16 and dQ = dS.K is the third MMA (A from TMEM).  Also writes delta for (B) and (C).
// (B) attn_tc_bwd_dkv_kernel — key-major, one CTA per 128-key tile: the same band seen from the keys.  S^T = K.Q^T,
//     dP^T = V.dO^T, then P^T and dS^T (lanes = keys) feed dV = P^T.dO and dK = dS^T.Q from tensor memory — gathered
//     per key like the other backward kernels: no atomics, deterministic.
// (C) attn_tc_bwd_edge_kernel — the duplicated edge keys (mhla.py:71-79): key N-1 also receives from the h rows clipped
//     on the left, key 0 from the h rows clipped on the right, with their multiplicities; one warp per (sequence, side)
//     adds those <= h terms to the two rows (B) wrote.  (A) adds the edge term of dQ for its own rows.
// =====================================================================================================================

// band of register k (column 32 w + k) for the row / key `r` = tile row 32 w + lane at sequence position `pos`:
// the other index is pos - h + (k - lane); valid iff 0 <= k - lane <= 2h and it lies in [0, N)
__device__ __forceinline__ void band_bounds(const TcParams& p, int pos, int lane, int& lo, int& hi) {
  lo = lane + max(0, p.h - pos);
  hi = lane + min(2 * p.h, p.N - 1 - pos + p.h);
  if (pos >= p.N) hi = lo - 1;
}

// packed bf16 pairs of this lane's NCH*32 values -> its own 16 NCH packed columns of the 96-column region at `region`,
// zeros over the rest of the region
template <int NCH>
__device__ __forceinline__ void store_packed_rows(uint32_t lane_base, uint32_t region, int w, const uint32_t (&pk)[NCH * 16]) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = pk[c * 16 + i];
    tmem_st16(lane_base + region + (uint32_t)(16 * w + 16 * c), r);
  }
  uint32_t z[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
  for (int t = 0; t < 6 - NCH; ++t) tmem_st16(lane_base + region + (uint32_t)(t < w ? 16 * t : 16 * (t + NCH)), z);
}

// dot product of two 64-element bf16 rows: `a` a swizzled shared-memory tile row (row index r), `b` a plain 128-byte row
__device__ __forceinline__ float dot64_swz(const uint8_t* tile, int r, const __nv_bfloat16* b) {
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 av = *reinterpret_cast<const uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4));
    const uint4 bv = *reinterpret_cast<const uint4*>(b + c * 8);
    const uint32_t aa[4] = {av.x, av.y, av.z, av.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      dot = fmaf(__uint_as_float(aa[t] << 16), __uint_as_float(bb[t] << 16), dot);
      dot = fmaf(__uint_as_float(aa[t] & 0xffff0000u), __uint_as_float(bb[t] & 0xffff0000u), dot);
    }
  }
  return dot;
}

struct TcBwdParams {
  TcParams f;
  const __nv_bfloat16* q;
  const __nv_bfloat16* o;     // [B,N,H,64] contiguous
  const __nv_bfloat16* dout;  // [B,N,H,64] contiguous
  const float* lse;           // [B,H,N]
  float* delta;               // [B,H,N]
  __nv_bfloat16* dq;
  __nv_bfloat16* dk;
  __nv_bfloat16* dv;          // same element strides as q / k / v
  float scale;
};

template <int NCH>
__global__ void __launch_bounds__(BQ, 2) attn_tc_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmq,
                                                               const __grid_constant__ CUtensorMap tmk,
                                                               const __grid_constant__ CUtensorMap tmv,
                                                               const __grid_constant__ CUtensorMap tmdo,
                                                               const TcBwdParams bp) {
  const TcParams& p = bp.f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t sQ = base, sDO = sQ + BQ * 128, sK = sDO + BQ * 128, sV = sK + (uint32_t)p.KT * 128;
  const uint32_t misc = sV + (uint32_t)p.KT * 128;
  const uint32_t bar_load = misc, bar_s = misc + 8, bar_dp = misc + 16, bar_dq = misc + 24, tmem_slot = misc + 32;
  uint8_t* g_misc = gen + (misc - base);
  __nv_bfloat16* s_edge = reinterpret_cast<__nv_bfloat16*>(g_misc + 64);   // k_0, k_{N-1}, v_0, v_{N-1}
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x % p.qtiles;
  const int bh = blockIdx.x / p.qtiles;
  const int hh = bh % p.H, b = bh / p.H;
  const int q0 = tile * BQ, kb = q0 - p.h;

  if (tid == 0) {
    prefetch_tmap(&tmq); prefetch_tmap(&tmk); prefetch_tmap(&tmv); prefetch_tmap(&tmdo);
    mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_dp, 1); mbar_init(bar_dq, 1);
    fence_barrier_init();
    mbar_expect_tx(bar_load, (uint32_t)(2 * BQ + 2 * p.KT) * 128u);
    tma_load_4d(sQ, &tmq, bar_load, 0, hh, q0, b);
    tma_load_4d(sK, &tmk, bar_load, 0, hh, kb, b);
    tma_load_4d(sDO, &tmdo, bar_load, 0, hh, q0, b);
    tma_load_4d(sV, &tmv, bar_load, 0, hh, kb, b);
  }
  if (warp == 1) tmem_alloc1(tmem_slot, 256);
  if (warp >= 2) {
    const int t = tid - 64;
    if (t < 32) {
      const int which = t >> 3, chunk = t & 7;
      const __nv_bfloat16* src = (which < 2 ? p.k : p.v) + (int64_t)b * p.sb + (int64_t)hh * p.sh +
                                 (int64_t)((which & 1) ? p.N - 1 : 0) * p.sn + chunk * 8;
      *reinterpret_cast<uint4*>(s_edge + which * 64 + chunk * 8) = *reinterpret_cast<const uint4*>(src);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(g_misc + 32);
  const uint32_t tmem_dq = tmem_base + (uint32_t)p.KT;

  if (tid == 0) {
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(BQ, p.KT, 0, 0);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k)
      umma_ss(tmem_base, make_smem_desc(sQ + k * 32u, 16u, 1024u), make_smem_desc(sK + k * 32u, 16u, 1024u), idesc,
              k > 0 ? 1u : 0u);
    umma_commit1(bar_s);
  }

  // ---- per-row scalars while the tensor core works: delta = dO . O, L (log2 domain), the duplicated-edge terms ----
  const int qi = q0 + tid;
  const bool live = qi < p.N;
  float delta = 0.f, L2 = CUDART_INF_F, pe = 0.f, dse = 0.f;
  int edge_row = -1;
  mbar_wait(bar_load, 0);
  if (live) {
    const __nv_bfloat16* orow = bp.o + ((int64_t)(b * p.N + qi) * p.H + hh) * HD;
    delta = dot64_swz(gen + (sDO - base), tid, orow);
    L2 = bp.lse[((int64_t)b * p.H + hh) * p.N + qi] * kLog2e;
    bp.delta[((int64_t)b * p.H + hh) * p.N + qi] = delta;
    const int s = max(0, qi - p.h), e = min(p.N, qi + p.h + 1);
    const int pad = p.W - (e - s);
    if (pad > 0) {
      edge_row = (s == 0) ? 1 : 0;
      const float dot = dot64_swz(gen + (sQ - base), tid, s_edge + edge_row * 64);
      pe = (float)pad * ex2f(dot * p.scale_log2 - L2);
      dse = pe * (dot64_swz(gen + (sDO - base), tid, s_edge + (2 + edge_row) * 64) - delta);
    }
  }
  int lo, hi;
  band_bounds(p, qi, lane, lo, hi);
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);

  // ---- P from S ----
  mbar_wait(bar_s, 0);
  tc_fence_after();
  float v[NCH * 32];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t r[32];
    tmem_ld32(lane_base + (uint32_t)(warp * 32 + c * 32), r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = c * 32 + i;
      const bool ok = k >= lo && k <= hi;
      v[k] = ok ? ex2f(__uint_as_float(r[i]) * p.scale_log2 - L2) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();   // every lane has read its S row: the accumulator columns can be overwritten
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(BQ, p.KT, 0, 0);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k)
      umma_ss(tmem_base, make_smem_desc(sDO + k * 32u, 16u, 1024u), make_smem_desc(sV + k * 32u, 16u, 1024u), idesc,
              k > 0 ? 1u : 0u);
    umma_commit1(bar_dp);
  }
  // ---- dS = P (dP - delta), packed, back to tensor memory ----
  mbar_wait(bar_dp, 0);
  tc_fence_after();
  uint32_t pk[NCH * 16];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t r[32];
    tmem_ld32(lane_base + (uint32_t)(warp * 32 + c * 32), r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d0 = v[c * 32 + 2 * i] * (__uint_as_float(r[2 * i]) - delta);
      const float d1 = v[c * 32 + 2 * i + 1] * (__uint_as_float(r[2 * i + 1]) - delta);
      pk[c * 16 + i] = pack_bf16x2(d0, d1);
    }
  }
  store_packed_rows<NCH>(lane_base, 0u, warp, pk);
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(BQ, HD, 0, 1);   // B = K: [keys][64], N contiguous
    for (int kk = 0; kk < p.KT / 16; ++kk)
      umma_ts(tmem_dq, tmem_base + (uint32_t)(kk * 8), make_smem_desc(sK + kk * 2048u, 8192u, 1024u), idesc, kk > 0 ? 1u : 0u);
    umma_commit1(bar_dq);
  }
  mbar_wait(bar_dq, 0);
  tc_fence_after();
  {
    __nv_bfloat16* drow = bp.dq + (int64_t)b * p.sb + (int64_t)(live ? qi : 0) * p.sn + (int64_t)hh * p.sh;
    const __nv_bfloat16* ke = s_edge + (edge_row < 0 ? 0 : edge_row) * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_dq + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), r);
      tmem_wait_ld();
      if (live) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float f[8];
#pragma unroll
          for (int e = 0; e < 8; ++e)
            f[e] = (__uint_as_float(r[g8 * 8 + e]) + dse * __bfloat162float(ke[c * 32 + g8 * 8 + e])) * bp.scale;
          store8(drow + c * 32 + g8 * 8, f);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc1(tmem_base, 256);
  }
}

template <int NCH>
__global__ void __launch_bounds__(BQ, 2) attn_tc_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmq,
                                                                const __grid_constant__ CUtensorMap tmk,
                                                                const __grid_constant__ CUtensorMap tmv,
                                                                const __grid_constant__ CUtensorMap tmdo,
                                                                const TcBwdParams bp) {
  const TcParams& p = bp.f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t sK = base, sV = sK + BQ * 128, sQ = sV + BQ * 128, sDO = sQ + (uint32_t)p.KT * 128;
  const uint32_t misc = sDO + (uint32_t)p.KT * 128;
  const uint32_t bar_load = misc, bar_s = misc + 8, bar_dp = misc + 16, bar_dv = misc + 24, bar_dk = misc + 32,
                 tmem_slot = misc + 40;
  uint8_t* g_misc = gen + (misc - base);
  float* s_L2 = reinterpret_cast<float*>(g_misc + 64);   // [KT] L_i in the log2 domain (+inf outside the sequence)
  float* s_dl = s_L2 + p.KT;                              // [KT] delta_i
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x % p.qtiles;
  const int bh = blockIdx.x / p.qtiles;
  const int hh = bh % p.H, b = bh / p.H;
  const int j0 = tile * BQ, qb = j0 - p.h;

  if (tid == 0) {
    prefetch_tmap(&tmq); prefetch_tmap(&tmk); prefetch_tmap(&tmv); prefetch_tmap(&tmdo);
    mbar_init(bar_load, 1); mbar_init(bar_s, 1); mbar_init(bar_dp, 1); mbar_init(bar_dv, 1); mbar_init(bar_dk, 1);
    fence_barrier_init();
    mbar_expect_tx(bar_load, (uint32_t)(2 * BQ + 2 * p.KT) * 128u);
    tma_load_4d(sK, &tmk, bar_load, 0, hh, j0, b);
    tma_load_4d(sQ, &tmq, bar_load, 0, hh, qb, b);
    tma_load_4d(sV, &tmv, bar_load, 0, hh, j0, b);
    tma_load_4d(sDO, &tmdo, bar_load, 0, hh, qb, b);
  }
  if (warp == 1) tmem_alloc1(tmem_slot, 256);
  for (int c = tid; c < p.KT; c += BQ) {
    const int i = qb + c;
    const bool in = i >= 0 && i < p.N;
    s_L2[c] = in ? bp.lse[((int64_t)b * p.H + hh) * p.N + i] * kLog2e : CUDART_INF_F;
    s_dl[c] = in ? bp.delta[((int64_t)b * p.H + hh) * p.N + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(g_misc + 40);

  if (tid == 0) {
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(BQ, p.KT, 0, 0);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k)
      umma_ss(tmem_base, make_smem_desc(sK + k * 32u, 16u, 1024u), make_smem_desc(sQ + k * 32u, 16u, 1024u), idesc,
              k > 0 ? 1u : 0u);
    umma_commit1(bar_s);
  }
  const int j = j0 + tid;
  const bool live = j < p.N;
  int lo, hi;
  band_bounds(p, j, lane, lo, hi);
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  const float* myL = s_L2 + warp * 32;
  const float* myD = s_dl + warp * 32;

  mbar_wait(bar_s, 0);
  tc_fence_after();
  float v[NCH * 32];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t r[32];
    tmem_ld32(lane_base + (uint32_t)(warp * 32 + c * 32), r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int k = c * 32 + i;
      const bool ok = k >= lo && k <= hi;
      v[k] = ok ? ex2f(__uint_as_float(r[i]) * p.scale_log2 - myL[k]) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(BQ, p.KT, 0, 0);
#pragma unroll
    for (int k = 0; k < HD / 16; ++k)
      umma_ss(tmem_base, make_smem_desc(sV + k * 32u, 16u, 1024u), make_smem_desc(sDO + k * 32u, 16u, 1024u), idesc,
              k > 0 ? 1u : 0u);
    umma_commit1(bar_dp);
  }
  mbar_wait(bar_dp, 0);
  tc_fence_after();
  uint32_t pp[NCH * 16], pd[NCH * 16];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    uint32_t r[32];
    tmem_ld32(lane_base + (uint32_t)(warp * 32 + c * 32), r);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int k = c * 32 + 2 * i;
      const float p0 = v[k], p1 = v[k + 1];
      pp[c * 16 + i] = pack_bf16x2(p0, p1);
      pd[c * 16 + i] = pack_bf16x2(p0 * (__uint_as_float(r[2 * i]) - myD[k]), p1 * (__uint_as_float(r[2 * i + 1]) - myD[k + 1]));
    }
  }
  store_packed_rows<NCH>(lane_base, 0u, warp, pp);     // P^T  -> columns [0, 96)
  store_packed_rows<NCH>(lane_base, 96u, warp, pd);    // dS^T -> columns [96, 192)
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(BQ, HD, 0, 1);
    for (int kk = 0; kk < p.KT / 16; ++kk)   // dV = P^T . dO -> columns [192, 256)
      umma_ts(tmem_base + 192u, tmem_base + (uint32_t)(kk * 8), make_smem_desc(sDO + kk * 2048u, 8192u, 1024u), idesc,
              kk > 0 ? 1u : 0u);
    umma_commit1(bar_dv);
    mbar_wait(bar_dv, 0);                     // P^T has been consumed: its columns take the dK accumulator
    tc_fence_after();
    for (int kk = 0; kk < p.KT / 16; ++kk)   // dK = dS^T . Q -> columns [0, 64)
      umma_ts(tmem_base, tmem_base + 96u + (uint32_t)(kk * 8), make_smem_desc(sQ + kk * 2048u, 8192u, 1024u), idesc,
              kk > 0 ? 1u : 0u);
    umma_commit1(bar_dk);
  }
  const int64_t roff = (int64_t)b * p.sb + (int64_t)(live ? j : 0) * p.sn + (int64_t)hh * p.sh;
  mbar_wait(bar_dv, 0);
  tc_fence_after();
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(lane_base + 192u + (uint32_t)(c * 32), r);
    tmem_wait_ld();
    if (live) {
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[g8 * 8 + e]);
        store8(bp.dv + roff + c * 32 + g8 * 8, f);
      }
    }
  }
  mbar_wait(bar_dk, 0);
  tc_fence_after();
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(lane_base + (uint32_t)(c * 32), r);
    tmem_wait_ld();
    if (live) {
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[g8 * 8 + e]) * bp.scale;
        store8(bp.dk + roff + c * 32 + g8 * 8, f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc1(tmem_base, 256);
  }
}

// (C) one warp per (b, h, side): side 0 = rows i < h (duplicates of key N-1), side 1 = rows i > N-1-h (duplicates of key 0)
__global__ void __launch_bounds__(256) attn_tc_bwd_edge_kernel(const TcBwdParams bp) {
  const TcParams& p = bp.f;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= (int64_t)p.B * p.H * 2) return;
  const int lane = threadIdx.x & 31;
  const int side = (int)(wid & 1);
  const int64_t bh = wid >> 1;
  const int hh = (int)(bh % p.H), b = (int)(bh / p.H);
  const int ekey = side == 0 ? p.N - 1 : 0;
  const int64_t seq = (int64_t)b * p.sb + (int64_t)hh * p.sh;
  auto ld2 = [&](const __nv_bfloat16* row, float& x, float& y) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(row + 2 * lane);
    x = __uint_as_float(u << 16);
    y = __uint_as_float(u & 0xffff0000u);
  };
  float k0, k1, v0, v1;
  ld2(p.k + seq + (int64_t)ekey * p.sn, k0, k1);
  ld2(p.v + seq + (int64_t)ekey * p.sn, v0, v1);
  float dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
  for (int t = 0; t < p.h; ++t) {
    const int i = side == 0 ? t : p.N - 1 - t;
    const int s = max(0, i - p.h), e = min(p.N, i + p.h + 1);
    const float pad = (float)(p.W - (e - s));
    float q0, q1, g0, g1;
    ld2(bp.q + seq + (int64_t)i * p.sn, q0, q1);
    ld2(bp.dout + ((int64_t)(b * p.N + i) * p.H + hh) * HD, g0, g1);
    float sd = q0 * k0 + q1 * k1, dp = g0 * v0 + g1 * v1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sd += __shfl_xor_sync(0xffffffffu, sd, o);
      dp += __shfl_xor_sync(0xffffffffu, dp, o);
    }
    const int64_t li = ((int64_t)b * p.H + hh) * p.N + i;
    const float pe = pad * ex2f(sd * p.scale_log2 - bp.lse[li] * kLog2e);
    const float ds = pe * (dp - bp.delta[li]) * bp.scale;
    dv0 = fmaf(pe, g0, dv0); dv1 = fmaf(pe, g1, dv1);
    dk0 = fmaf(ds, q0, dk0); dk1 = fmaf(ds, q1, dk1);
  }
  __nv_bfloat16* dkr = bp.dk + seq + (int64_t)ekey * p.sn + 2 * lane;
  __nv_bfloat16* dvr = bp.dv + seq + (int64_t)ekey * p.sn + 2 * lane;
  float a0, a1;
  ld2(bp.dk + seq + (int64_t)ekey * p.sn, a0, a1);
  *reinterpret_cast<uint32_t*>(dkr) = pack_bf16x2(a0 + dk0, a1 + dk1);
  ld2(bp.dv + seq + (int64_t)ekey * p.sn, a0, a1);
  *reinterpret_cast<uint32_t*>(dvr) = pack_bf16x2(a0 + dv0, a1 + dv1);
}

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

// [B][N][H][64] bf16 view with element strides (sb, sn, sh); box = 64 x 1 x rows x 1
int make_map(CUtensorMap* tm, const void* ptr, int B, int H, int N, int64_t sb, int64_t sn, int64_t shh, int rows) {
  cuuint64_t gdim[4] = {(cuuint64_t)HD, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)shh * 2, (cuuint64_t)sn * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {(cuuint32_t)HD, 1, (cuuint32_t)rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = encode_fn()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attn_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return FAVIT_ERR_CUDA;
  }
  return FAVIT_OK;
}

}  // namespace

bool attn_tc_applicable(int hd, int window, int N, favit_dtype dtype, const uint8_t* mask, const void* q, const void* k,
                        const void* v, int64_t sb, int64_t sn, int64_t shh) {
  auto al = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
  return dtype == FAVIT_BF16 && mask == nullptr && hd == HD && (window & 1) && window >= 17 && window <= 65 && N >= window &&
         al(q) && al(k) && al(v) && sb % 8 == 0 && sn % 8 == 0 && shh % 8 == 0 && encode_fn() != nullptr;
}

int attn_tc_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int window,
                float scale, int64_t sb, int64_t sn, int64_t shh, cudaStream_t st) {
  TcParams p;
  p.B = B; p.H = H; p.N = N; p.W = window; p.h = window / 2;
  p.KT = (BQ + 2 * p.h + 15) / 16 * 16;
  p.qtiles = ceil_div(N, BQ);
  p.scale_log2 = scale * kLog2e;
  p.k = (const __nv_bfloat16*)k;
  p.v = (const __nv_bfloat16*)v;
  p.sb = sb; p.sn = sn; p.sh = shh;
  CUtensorMap tq, tk, tv;
  if (int rc = make_map(&tq, q, B, H, N, sb, sn, shh, BQ)) return rc;
  if (int rc = make_map(&tk, k, B, H, N, sb, sn, shh, p.KT)) return rc;
  if (int rc = make_map(&tv, v, B, H, N, sb, sn, shh, p.KT)) return rc;
  // at least 78 KB per CTA: two CTAs per SM, which is what tensor memory (2 x 256 columns) allows — a third one would sit
  // in tcgen05.alloc's retry loop holding shared memory and a barrier's worth of warps
  const size_t smem = std::max<size_t>((size_t)(BQ + 2 * p.KT) * 128 + 64 + 4 * 128 + 1024, 78 * 1024);
  const int64_t grid = (int64_t)B * H * p.qtiles;
  FAVIT_CHECK_ARG(grid < INT32_MAX, "attn_tc_fwd: grid too large");
  const int nch = (32 + 2 * p.h + 31) / 32;     // S chunks per warp: 2 (W <= 33) or 3
  const int npch = (p.KT / 2 + 31) / 32;         // packed-P store chunks: 3 for KT in (128, 192]
  static bool configured = false;
  if (!configured) {
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<3, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  if (npch != 3 || nch < 2 || nch > 3) {
    set_error("attn_tc_fwd: window %d unsupported (internal)", window);
    return FAVIT_ERR_UNSUPPORTED;
  }
  if (nch == 2)
    attn_tc_fwd_kernel<2, 3><<<(unsigned)grid, BQ, smem, st>>>(tq, tk, tv, (__nv_bfloat16*)out, lse, p);
  else
    attn_tc_fwd_kernel<3, 3><<<(unsigned)grid, BQ, smem, st>>>(tq, tk, tv, (__nv_bfloat16*)out, lse, p);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

int attn_tc_bwd(const void* q, const void* k, const void* v, const void* out, const float* lse, const void* dout, void* dq,
                void* dk, void* dv, float* delta, int B, int H, int N, int window, float scale, int64_t sb, int64_t sn,
                int64_t shh, cudaStream_t st) {
  TcBwdParams bp;
  TcParams& p = bp.f;
  p.B = B; p.H = H; p.N = N; p.W = window; p.h = window / 2;
  p.KT = (BQ + 2 * p.h + 15) / 16 * 16;
  p.qtiles = ceil_div(N, BQ);
  p.scale_log2 = scale * kLog2e;
  p.k = (const __nv_bfloat16*)k;
  p.v = (const __nv_bfloat16*)v;
  p.sb = sb; p.sn = sn; p.sh = shh;
  bp.q = (const __nv_bfloat16*)q;
  bp.o = (const __nv_bfloat16*)out;
  bp.dout = (const __nv_bfloat16*)dout;
  bp.lse = lse;
  bp.delta = delta;
  bp.dq = (__nv_bfloat16*)dq; bp.dk = (__nv_bfloat16*)dk; bp.dv = (__nv_bfloat16*)dv;
  bp.scale = scale;
  const int64_t dsb = (int64_t)N * H * HD, dsn = (int64_t)H * HD;
  CUtensorMap tq128, tk128, tv128, td128, tqK, tkK, tvK, tdK;
  if (int rc = make_map(&tq128, q, B, H, N, sb, sn, shh, BQ)) return rc;
  if (int rc = make_map(&tk128, k, B, H, N, sb, sn, shh, BQ)) return rc;
  if (int rc = make_map(&tv128, v, B, H, N, sb, sn, shh, BQ)) return rc;
  if (int rc = make_map(&td128, dout, B, H, N, dsb, dsn, HD, BQ)) return rc;
  if (int rc = make_map(&tqK, q, B, H, N, sb, sn, shh, p.KT)) return rc;
  if (int rc = make_map(&tkK, k, B, H, N, sb, sn, shh, p.KT)) return rc;
  if (int rc = make_map(&tvK, v, B, H, N, sb, sn, shh, p.KT)) return rc;
  if (int rc = make_map(&tdK, dout, B, H, N, dsb, dsn, HD, p.KT)) return rc;
  const size_t smem = std::max<size_t>((size_t)(2 * BQ + 2 * p.KT) * 128 + 64 + 4 * 128 + 2 * (size_t)p.KT * 4 + 1024, 78 * 1024);
  const int64_t grid = (int64_t)B * H * p.qtiles;
  FAVIT_CHECK_ARG(grid < INT32_MAX, "attn_tc_bwd: grid too large");
  const int nch = (32 + 2 * p.h + 31) / 32;
  static bool configured = false;
  if (!configured) {
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_dq_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_dq_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_dkv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_bwd_dkv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    configured = true;
  }
  if (nch == 2) {
    attn_tc_bwd_dq_kernel<2><<<(unsigned)grid, BQ, smem, st>>>(tq128, tkK, tvK, td128, bp);
    FAVIT_CHECK_LAUNCH();
    attn_tc_bwd_dkv_kernel<2><<<(unsigned)grid, BQ, smem, st>>>(tqK, tk128, tv128, tdK, bp);
    FAVIT_CHECK_LAUNCH();
  } else if (nch == 3) {
    attn_tc_bwd_dq_kernel<3><<<(unsigned)grid, BQ, smem, st>>>(tq128, tkK, tvK, td128, bp);
    FAVIT_CHECK_LAUNCH();
    attn_tc_bwd_dkv_kernel<3><<<(unsigned)grid, BQ, smem, st>>>(tqK, tk128, tv128, tdK, bp);
    FAVIT_CHECK_LAUNCH();
  } else {
    set_error("attn_tc_bwd: window %d unsupported (internal)", window);
    return FAVIT_ERR_UNSUPPORTED;
  }
  const int64_t warps = (int64_t)B * H * 2;
  attn_tc_bwd_edge_kernel<<<(unsigned)ceil_div64(warps, 8), 256, 0, st>>>(bp);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

}  // namespace favit
