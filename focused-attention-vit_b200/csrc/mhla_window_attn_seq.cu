// MHLA windowed attention, whole-sequence tiles (bf16, head_dim 64, no mask, window <= 15, N <= 400), sm_100a.
//
// Same math and reference span as mhla_window_attn.cu (/root/reference/models/mhla.py:109-154: banded softmax with the
// duplicated-edge multiplicities of mhla.py:72-79) and the same mma.sync tile arithmetic as mhla_window_attn_mma.cu.
// What changes is who moves the bytes.  There, every warp copies its own 16 queries + 32 key/value rows with per-lane
// cp.async (5x read amplification out of L2, ~500 address instructions per tile, one exposed latency per warp).  Here a
// CTA owns G whole (image, head) sequences: one thread issues a TMA box per operand ([N x 64] bf16 out of the packed
// qkv tensor, SWIZZLE_128B, which is exactly the ldmatrix-conflict-free layout the tiles use), every row is fetched
// once, and the warps (one per 16-row tile) read their band rows and the two edge rows (key N-1 / key 0) in place
// through a per-lane slot -> row map.  Two CTAs per SM overlap one CTA's loads with the other's arithmetic.
// Backward is ONE kernel: a query-major phase (delta, dQ) and a key-major phase (dK, dV) over the same four resident
// tiles (Q, K, V, dO), so Q/K/V/dO are read once instead of twice and delta never goes through global memory.
// The op is HBM-bound (AI = W/2 FLOP/B): algorithmic bytes fwd = 4*B*N*D*2, bwd = 8*B*N*D*2.
#include <cuda.h>

#include <algorithm>
#include <mutex>

#include "attn_seq_common.cuh"

namespace favit {
namespace {

using namespace attn;
using namespace seqk;


struct SeqParams {
  Shape sh;
  int G;           // (image, head) sequences per CTA
  int alloc_rows;  // rows of one shared-memory tile: tiles * 16
  int rows_box;    // rows per TMA box (multiple of 8, <= 256)
  int nbox;        // boxes per operand
  int64_t pairs;   // B * H
  int cs_shared;   // backward: column sums through the CTA's shared-memory accumulator (few heads: few, hot addresses)
  int ahead;       // CTAs: the operands of CTA blockIdx + ahead are prefetched into L2 while this one computes (0: off)
};

// One thread: barrier init + every TMA box of the CTA's sequences.  NOPS operand tiles per sequence.
template <int NOPS>
__device__ __forceinline__ void issue_loads(const CUtensorMap* const (&tm)[NOPS], uint8_t* tiles, size_t pair_bytes,
                                            size_t tile_bytes, uint32_t bar, const SeqParams& p, int64_t pair0) {
  ptx::mbar_init(bar, 1);
  ptx::fence_barrier_init();
  ptx::fence_proxy_async_smem();
  int npairs = 0;
  for (int g = 0; g < p.G; ++g) npairs += (pair0 + g < p.pairs) ? 1 : 0;
  ptx::mbar_expect_tx(bar, (uint32_t)npairs * NOPS * p.nbox * p.rows_box * kRowBytes);
  for (int g = 0; g < p.G; ++g) {
    const int64_t pr = pair0 + g;
    if (pr >= p.pairs) break;
    const int b = (int)(pr / p.sh.H), h = (int)(pr % p.sh.H);
#pragma unroll
    for (int t = 0; t < NOPS; ++t)
      for (int x = 0; x < p.nbox; ++x)
        tma_load_4d(smem_u32(tiles + g * pair_bytes + t * tile_bytes + (size_t)x * p.rows_box * kRowBytes), tm[t], bar, 0,
                    h, x * p.rows_box, b);
  }
  // The CTA that will take this one's place on the SM (blockIdx + the number of resident CTAs) finds its operands in
  // L2: nobody waits for these reads, so HBM keeps streaming while every resident CTA is in its arithmetic phase.
  if (p.ahead > 0) {
    const int64_t next0 = pair0 + (int64_t)p.ahead * p.G;
    for (int g = 0; g < p.G; ++g) {
      const int64_t pr = next0 + g;
      if (pr >= p.pairs) break;
      const int b = (int)(pr / p.sh.H), h = (int)(pr % p.sh.H);
#pragma unroll
      for (int t = 0; t < NOPS; ++t)
        for (int x = 0; x < p.nbox; ++x) tma_prefetch_4d(tm[t], 0, h, x * p.rows_box, b);
    }
  }
}

// rows [nbox * rows_box, alloc_rows) of every tile are not written by TMA: make them finite (zero)
__device__ __forceinline__ void zero_tail_rows(uint8_t* tiles, int ntiles_total, size_t tile_bytes, const SeqParams& p) {
  const int r0 = p.nbox * p.rows_box;
  const int n16 = (p.alloc_rows - r0) * (kRowBytes / 16);
  for (int t = 0; t < ntiles_total; ++t)
    for (int i = threadIdx.x; i < n16; i += blockDim.x)
      *reinterpret_cast<uint4*>(tiles + t * tile_bytes + (size_t)r0 * kRowBytes + i * 16) = make_uint4(0, 0, 0, 0);
}

// =================================================================================================================
// forward
// =================================================================================================================
template <int NT>
__global__ void __launch_bounds__(kMaxThreads) attn_seq_fwd_kernel(const __grid_constant__ CUtensorMap tmq,
                                                                  const __grid_constant__ CUtensorMap tmk,
                                                                  const __grid_constant__ CUtensorMap tmv,
                                                                  __nv_bfloat16* __restrict__ out,
                                                                  float* __restrict__ lse, SeqParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // the 128B swizzle is keyed on address bits
  const Shape& sh = p.sh;
  const size_t tile_bytes = (size_t)p.alloc_rows * kRowBytes, pair_bytes = 3 * tile_bytes;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + p.G * pair_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pair0 = (int64_t)blockIdx.x * p.G;
  if (threadIdx.x == 0) {
    const CUtensorMap* const tm[3] = {&tmq, &tmk, &tmv};
    issue_loads<3>(tm, smem, pair_bytes, tile_bytes, smem_u32(bar), p, pair0);
  }
  zero_tail_rows(smem, 3 * p.G, tile_bytes, p);
  __syncthreads();
  ptx::mbar_wait(smem_u32(bar), 0);

  const int g = warp / sh.tiles, qt = warp - g * sh.tiles;
  const int64_t pr = pair0 + g;
  if (pr >= p.pairs) return;
  const int b = (int)(pr / sh.H), h = (int)(pr % sh.H);
  const uint8_t* sQ = smem + g * pair_bytes;
  const uint8_t* sK = sQ + tile_bytes;
  const uint8_t* sV = sK + tile_bytes;
  const int N = sh.N, i0 = qt * 16;
  uint8_t* sQt = const_cast<uint8_t*>(sQ) + (size_t)i0 * kRowBytes;  // own query tile, later the output staging tile
  const KeySlots ks = key_slots(i0, N, sh.W);
  int krow[NT / 2], vrow[NT / 2];
#pragma unroll
  for (int x = 0; x < NT / 2; ++x) {
    krow[x] = max(ks.key(16 * x + (lane & 7) + 8 * (lane >> 4)), 0);        // unused slots read row 0; their P is 0
    vrow[x] = max(ks.key(16 * x + (lane & 7) + 8 * ((lane >> 3) & 1)), 0);
  }

  float s[NT][4];
  scores_rows<NT>(sQt, sK, krow, lane, s);

  const int r0 = lane >> 2;
  const RowSlots w0 = row_slots(ks, min(i0 + r0, N - 1), N, sh.W), w1 = row_slots(ks, min(i0 + r0 + 8, N - 1), N, sh.W);
  float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int slot = nt * 8 + (lane & 3) * 2 + (e & 1);
      const float val = fmaf(s[nt][e], sh.scale_log2, (e < 2 ? w0 : w1).bias(slot));  // -inf outside the window
      s[nt][e] = val;
      if (e < 2) mx0 = fmaxf(mx0, val); else mx1 = fmaxf(mx1, val);
    }
  mx0 = quad_max(mx0);
  mx1 = quad_max(mx1);
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float pe = exp2f(s[nt][e] - (e < 2 ? mx0 : mx1));
      s[nt][e] = pe;
      if (e < 2) sum0 += pe; else sum1 += pe;
    }
  sum0 = quad_sum(sum0);
  sum1 = quad_sum(sum1);

  float o[HD / 8][4];
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nd][e] = 0.f;
  pv_rows<NT>(s, sV, vrow, lane, o);

  __syncwarp();  // the query rows of a tile are read by its own warp only: reuse them as staging
  stage_acc<HD>(sQt, o, 1.f / sum0, 1.f / sum1, lane);
  __syncwarp();
  store_rows<HD>(sQt, lane, [&](int r) {
    return (i0 + r < N) ? out + (((int64_t)b * N + i0 + r) * sh.H + h) * HD : nullptr;
  });
  if ((lane & 3) == 0) {
    float* l = lse + ((int64_t)b * sh.H + h) * N;
    if (i0 + r0 < N) l[i0 + r0] = (mx0 + log2f(sum0)) * kLn2;
    if (i0 + r0 + 8 < N) l[i0 + r0 + 8] = (mx1 + log2f(sum1)) * kLn2;
  }
}

// =================================================================================================================
// backward: phase A (query tiles: delta, P, dS, dQ), barrier, P / dS parked as bf16 band blocks, barrier,
//           phase B (key tiles: dV = P^T.dO, dK = scale.dS^T.Q straight from the parked blocks)
// =================================================================================================================
template <int NT>
__global__ void __launch_bounds__(kMaxThreads) attn_seq_bwd_kernel(
    const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
    const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
    const __nv_bfloat16* __restrict__ o, const float* __restrict__ lse, __nv_bfloat16* __restrict__ dq,
    __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, SeqParams p) {
  static_assert(NT == 4, "32 aligned key slots per query tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // the 128B swizzle is keyed on address bits
  const Shape& sh = p.sh;
  const size_t tile_bytes = (size_t)p.alloc_rows * kRowBytes, pair_bytes = 4 * tile_bytes;
  float* sLall = reinterpret_cast<float*>(smem + p.G * pair_bytes);  // [G][alloc_rows] log2-domain LSE (+inf past N)
  float* sDall = sLall + p.G * p.alloc_rows;                          // [G][alloc_rows] delta
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sDall + p.G * p.alloc_rows);
  uint8_t* zero16 = reinterpret_cast<uint8_t*>(bar + 2);              // 16 zero bytes: the absent 8 x 8 blocks
  // [G][3 * 64] column sums of dQ | dK | dV of this CTA's sequences: warps add here (shared-memory atomics), the CTA
  // adds each column to global memory once.  With one global atomic per column per WARP, 40 000 warps queue on 2 304
  // addresses and that queue, not HBM, sets the kernel time (measured: ~100 us of fixed cost at every size).
  float* sCS = reinterpret_cast<float*>(zero16 + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pair0 = (int64_t)blockIdx.x * p.G;
  if (threadIdx.x == 0) {
    const CUtensorMap* const tm[4] = {&tmq, &tmk, &tmv, &tmdo};
    issue_loads<4>(tm, smem, pair_bytes, tile_bytes, smem_u32(bar), p, pair0);
  }
  if (threadIdx.x < 4) reinterpret_cast<uint32_t*>(zero16)[threadIdx.x] = 0u;
  for (int i = threadIdx.x; i < p.G * 3 * HD; i += blockDim.x) sCS[i] = 0.f;
  zero_tail_rows(smem, 4 * p.G, tile_bytes, p);
  const int tiles = sh.tiles;
  const int g = warp / tiles, tt = warp - g * tiles;
  const int64_t pr = pair0 + g;
  const bool live = pr < p.pairs;
  const int b = live ? (int)(pr / sh.H) : 0, h = live ? (int)(pr % sh.H) : 0;
  // O (for delta) and the LSE come straight from global: fetch them while the TMA boxes are in flight
  uint4 o_pre[4];
  float lse_pre = 0.f;
  {
    const int i0 = tt * 16;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = (lane >> 3) + 4 * j, c = lane & 7;
      o_pre[j] = make_uint4(0, 0, 0, 0);
      if (live && i0 + r < sh.N)
        o_pre[j] = *reinterpret_cast<const uint4*>(o + (((int64_t)b * sh.N + i0 + r) * sh.H + h) * HD + c * 8);
    }
    if (live && lane < 16 && i0 + lane < sh.N) lse_pre = lse[((int64_t)b * sh.H + h) * sh.N + i0 + lane];
  }
  __syncthreads();
  ptx::mbar_wait(smem_u32(bar), 0);

  uint8_t* sQ = smem + g * pair_bytes;
  uint8_t* sK = sQ + tile_bytes;    // after phase A: the parked P blocks [tiles][16][32], then the dS blocks
  uint8_t* sV = sK + tile_bytes;    // after phase A: output staging
  uint8_t* sdO = sV + tile_bytes;
  float* sL = sLall + g * p.alloc_rows;
  float* sD = sDall + g * p.alloc_rows;
  const int N = sh.N, t0 = tt * 16;
  const int64_t base = (int64_t)b * sh.sb + (int64_t)h * sh.sh;
  const int r0 = lane >> 2;
  uint32_t p_pk[NT][2], ds_pk[NT][2];  // this tile's P and dS, bf16 pairs in accumulator-fragment order

  // ---------------------------------------------------------------- phase A: queries t0 .. t0+15
  if (live) {
    const int i0 = t0;
    const uint8_t* sQt = sQ + (size_t)i0 * kRowBytes;
    const uint8_t* sdOt = sdO + (size_t)i0 * kRowBytes;
    // delta_i = dO_i . O_i: O straight from global (16-byte coalesced), dO from the resident tile
    {
      if (lane < 16) sL[i0 + lane] = (i0 + lane < N) ? lse_pre * kLog2e : CUDART_INF_F;  // +inf -> P = 0
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = (lane >> 3) + 4 * j, c = lane & 7;
        float part = 0.f;
        if (i0 + r < N) {
          const uint4 ov = o_pre[j];
          const uint4 dv4 = *reinterpret_cast<const uint4*>(sdOt + tile_off<HD>(r, c));
          part = bf16x2_dot(ov.x, dv4.x) + bf16x2_dot(ov.y, dv4.y) + bf16x2_dot(ov.z, dv4.z) + bf16x2_dot(ov.w, dv4.w);
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        if (c == 0) sD[i0 + r] = part;
      }
      __syncwarp();
    }
    const AlSlots ks = al_slots(i0, N, sh.W);
    int krow[NT / 2], vrow[NT / 2];
#pragma unroll
    for (int x = 0; x < NT / 2; ++x) {
      krow[x] = max(ks.key(16 * x + (lane & 7) + 8 * (lane >> 4)), 0);        // unused slots read row 0; their P is 0
      vrow[x] = max(ks.key(16 * x + (lane & 7) + 8 * ((lane >> 3) & 1)), 0);
    }
    float s[NT][4], dp[NT][4];
    scores_rows<NT>(sQt, sK, krow, lane, s);
    scores_rows<NT>(sdOt, sV, krow, lane, dp);
    const bool ok0 = i0 + r0 < N, ok1 = i0 + r0 + 8 < N;
    const RowSlots w0 = al_row_slots(ks, min(i0 + r0, N - 1), N, sh.W);
    const RowSlots w1 = al_row_slots(ks, min(i0 + r0 + 8, N - 1), N, sh.W);
    const float L0 = sL[i0 + r0], L1 = sL[i0 + r0 + 8], d0 = sD[i0 + r0], d1 = sD[i0 + r0 + 8];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int slot = nt * 8 + (lane & 3) * 2 + (e & 1);
        const float pe = exp2f(fmaf(s[nt][e], sh.scale_log2, (e < 2 ? w0 : w1).bias(slot)) - (e < 2 ? L0 : L1));
        s[nt][e] = pe * (dp[nt][e] - (e < 2 ? d0 : d1));
        dp[nt][e] = pe;
      }
      p_pk[nt][0] = pack_bf16x2(dp[nt][0], dp[nt][1]);
      p_pk[nt][1] = pack_bf16x2(dp[nt][2], dp[nt][3]);
      ds_pk[nt][0] = pack_bf16x2(s[nt][0], s[nt][1]);
      ds_pk[nt][1] = pack_bf16x2(s[nt][2], s[nt][3]);
    }
    float acc[HD / 8][4];
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {  // dQ = scale . dS . K
      const uint32_t a[4] = {ds_pk[2 * kk][0], ds_pk[2 * kk][1], ds_pk[2 * kk + 1][0], ds_pk[2 * kk + 1][1]};
      mma_rows(a, sK, vrow[kk], lane, acc);
    }
    // dQ leaves from the accumulator fragments (the Q rows are operands of the neighbours' phase B): a quad writes
    // 16 contiguous bytes, two n-steps complete a 32-byte sector in L2
    __nv_bfloat16* dq0 = dq + base + (int64_t)(i0 + r0) * sh.sn + (lane & 3) * 2;
    __nv_bfloat16* dq1 = dq0 + 8 * sh.sn;
    uint32_t v0[HD / 8], v1[HD / 8];
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd) {
      v0[nd] = ok0 ? pack_bf16x2(acc[nd][0] * sh.scale, acc[nd][1] * sh.scale) : 0u;  // rows past N: exact zeros
      v1[nd] = ok1 ? pack_bf16x2(acc[nd][2] * sh.scale, acc[nd][3] * sh.scale) : 0u;
      if (ok0) *reinterpret_cast<uint32_t*>(dq0 + nd * 8) = v0[nd];
      if (ok1) *reinterpret_cast<uint32_t*>(dq1 + nd * 8) = v1[nd];
    }
    if (sh.colsum)  // the q rows of the qkv bias gradient
      colsum_frags(v0, v1, lane, p.cs_shared ? sCS + g * 3 * HD : sh.colsum + h * HD, p.cs_shared);
  }
  __syncthreads();  // nobody reads K / V band rows any more
  if (live) {       // park P in the K tile's first half, dS in its second half
    uint8_t* pP = sK + (size_t)tt * kBlkBytes;
    uint8_t* pS = pP + (size_t)tiles * kBlkBytes;
    const int sub = (lane & 3) * 4;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      *reinterpret_cast<uint32_t*>(pP + tile_off<32>(r0, nt) + sub) = p_pk[nt][0];
      *reinterpret_cast<uint32_t*>(pP + tile_off<32>(r0 + 8, nt) + sub) = p_pk[nt][1];
      *reinterpret_cast<uint32_t*>(pS + tile_off<32>(r0, nt) + sub) = ds_pk[nt][0];
      *reinterpret_cast<uint32_t*>(pS + tile_off<32>(r0 + 8, nt) + sub) = ds_pk[nt][1];
    }
  }
  __syncthreads();

  // ---------------------------------------------------------------- phase B: keys t0 .. t0+15
  if (live) {
    const int j0 = t0;
    const uint8_t* sP = sK;
    const uint8_t* sS = sK + (size_t)tiles * kBlkBytes;
    const uint32_t zaddr = smem_u32(zero16);
    // ldmatrix.x4.trans row addresses of the A fragments P^T / dS^T (offsets relative to sP / sS):
    //   k-step 0: queries [j0-8, j0+8) = tile t-1 rows 8-15, tile t rows 0-7; k-step 1: [j0+8, j0+24) = tile t rows
    //   8-15, tile t+1 rows 0-7.  Matrix m of the x4 = lanes 8m..8m+7: (k half, key half) = (m >> 1, m & 1).
    const int mi = lane >> 3, r8 = lane & 7;
    int offA[2];
    {
      // k-step 0: m0 (t-1, rows 8.., chunk 3)  m1 zero                 m2 (t, rows 0.., chunk 1)  m3 (t, rows 0.., chunk 2)
      // k-step 1: m0 (t, rows 8.., chunk 1)    m1 (t, rows 8.., chunk 2) m2 zero                  m3 (t+1, rows 0.., chunk 0)
      const int q0 = (mi == 0) ? tt - 1 : tt, rr0 = (mi == 0) ? 8 + r8 : r8, c0 = (mi == 0) ? 3 : mi - 1;
      offA[0] = (mi == 1 || q0 < 0) ? -1 : q0 * kBlkBytes + (int)tile_off<32>(rr0, c0);
      const int q1 = (mi == 3) ? tt + 1 : tt, rr1 = (mi == 3) ? r8 : 8 + r8, c1 = (mi == 3) ? 0 : mi + 1;
      offA[1] = (mi == 2 || q1 >= tiles) ? -1 : q1 * kBlkBytes + (int)tile_off<32>(rr1, c1);
    }
    int brow[2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
      brow[kk] = min(max(j0 - 8 + 16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1), 0), p.alloc_rows - 1);
    const int ja = j0 + r0, jb = j0 + r0 + 8;
    const bool okA = ja < N, okB = jb < N;
    // edge keys: key N-1 (last tile) collects slot 0 of queries 0..15, key 0 (first tile) collects slot 31 of the last
    // 16 queries (late queries are within h of the end); rows that duplicate nothing parked zeros there
    const int tlast = (N - 1) >> 4;
    const bool edgeA = tt == tlast, edgeB = tt == 0;
    const int qb = max(N - 16, 0);
    const int k0 = (lane & 3) * 2;
    auto val = [&](const uint8_t* blocks, int q, int slot) -> uint32_t {  // bf16 bits parked for query q, slot `slot`
      return *reinterpret_cast<const uint16_t*>(blocks + (size_t)(q >> 4) * kBlkBytes + tile_off<32>(q & 15, slot >> 3) +
                                                (slot & 7) * 2);
    };
    auto edge_frag = [&](const uint8_t* blocks, bool keyA, uint32_t (&a)[4]) {
      a[0] = a[1] = a[2] = a[3] = 0u;
      const int R = keyA ? (N - 1) & 15 : 0;       // row of the edge key in this tile
      const int q = keyA ? k0 : qb + k0, slot = keyA ? 0 : 31;
      if (r0 == (R & 7)) {  // A[R][k] = value(query q_first + k, slot)
        const uint32_t lo = val(blocks, q, slot) | (val(blocks, q + 1, slot) << 16);
        const uint32_t hi = val(blocks, q + 8, slot) | (val(blocks, q + 9, slot) << 16);
        if (R < 8) { a[0] = lo; a[2] = hi; } else { a[1] = lo; a[3] = hi; }
      }
    };
    const int erowA = (lane & 7) + 8 * ((lane >> 3) & 1), erowB = qb + erowA;  // B rows of the edge steps
    uint8_t* stage = sV + (size_t)j0 * kRowBytes;  // the V rows of this tile: free since the barrier
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {  // 0: dV = P^T . dO     1: dK = scale . dS^T . Q
      const uint8_t* blocks = pass == 0 ? sP : sS;
      const uint8_t* rows = pass == 0 ? sdO : sQ;
      const uint32_t bbase = smem_u32(blocks);
      float acc[HD / 8][4];
#pragma unroll
      for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a[4];
        ldsm_x4_trans(offA[kk] >= 0 ? bbase + (uint32_t)offA[kk] : zaddr, a);
        mma_rows(a, rows, brow[kk], lane, acc);
      }
      if (edgeA) {  // warp-uniform
        uint32_t a[4];
        edge_frag(blocks, true, a);
        mma_rows(a, rows, erowA, lane, acc);
      }
      if (edgeB) {
        uint32_t a[4];
        edge_frag(blocks, false, a);
        mma_rows(a, rows, erowB, lane, acc);
      }
      const float m = pass == 0 ? 1.f : sh.scale;
      __syncwarp();
      stage_acc<HD>(stage, acc, okA ? m : 0.f, okB ? m : 0.f, lane);  // rows past N must stay exact zeros (column sums)
      __syncwarp();
      __nv_bfloat16* dst = pass == 0 ? dv : dk;
      store_rows<HD>(stage, lane, [&](int r) { return (j0 + r < N) ? dst + base + (int64_t)(j0 + r) * sh.sn : nullptr; });
      if (sh.colsum)
        colsum_staged(stage, lane, p.cs_shared ? sCS + (g * 3 + (pass == 0 ? 2 : 1)) * HD
                                               : sh.colsum + ((pass == 0 ? 2 : 1) * sh.H + h) * HD, p.cs_shared);
    }
  }
  if (sh.colsum && p.cs_shared) {  // one global atomic per column per CTA
    __syncthreads();
    for (int i = threadIdx.x; i < p.G * 3 * HD; i += blockDim.x) {
      const int g2 = i / (3 * HD), c = i - g2 * 3 * HD;
      const int64_t pr2 = pair0 + g2;
      if (pr2 < p.pairs) atomicAdd(sh.colsum + ((c / HD) * sh.H + (int)(pr2 % sh.H)) * HD + (c % HD), sCS[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

SeqParams make_params(int B, int H, int N, int window, float scale, int64_t sb, int64_t sn, int64_t shh, float* colsum,
                      int nops) {
  SeqParams p;
  const int tiles = ceil_div(N, 16);
  p.sh = Shape{B, H, N, window, sb, sn, shh, scale * kLog2e, scale, tiles, colsum};
  p.nbox = ceil_div(N, 256);
  p.rows_box = ((ceil_div(N, p.nbox) + 7) / 8) * 8;
  p.alloc_rows = std::max(tiles * 16, p.nbox * p.rows_box);
  p.pairs = (int64_t)B * H;
  const size_t per_pair = (size_t)nops * p.alloc_rows * kRowBytes + (nops == 4 ? 2 * p.alloc_rows * 4 : 0);
  int G = std::max(1, 4 / tiles);  // at least four warps per CTA; short sequences still get many small CTAs per SM
  G = std::min(G, std::max(1, (int)((100 * 1024) / per_pair)));
  G = std::min(G, 8);
  p.G = (int)std::min<int64_t>(G, p.pairs);
  // 3*H*64 global addresses take one reduction per column per warp; measured crossover of the two schemes: H = 6 / 12
  p.cs_shared = H <= 8 ? 1 : 0;
  p.ahead = 0;
  return p;
}

// L2 prefetch distance of the forward kernel, in CTAs: half of the CTAs resident at once (measured best of 0 / 0.5 / 1 /
// 1.5 / 2 x resident at ViT-B, ViT-S and CIFAR shapes: -5 % time; profiles/r2_attn_seq_probe.txt), 0 when the grid fits
// in one wave.  The backward kernel does not prefetch (same sweep: +3 %; its CTAs wait on arithmetic, not on loads).
// FAVIT_SEQ_AHEAD = n overrides both (tuning only; n < 0: -n percent of the resident count).
int prefetch_distance(const void* kernel, int threads, size_t smem, unsigned grid, int default_pct) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) {
    (void)cudaGetLastError();
    return 0;
  }
  const int resident = occ * num_sms();
  int ahead = resident * default_pct / 100;
  if (const char* e = getenv("FAVIT_SEQ_AHEAD")) ahead = atoi(e) < 0 ? resident * -atoi(e) / 100 : atoi(e);
  return (unsigned)ahead < grid ? ahead : 0;
}

}  // namespace

bool attn_seq_applicable(int hd, int window, int N, favit_dtype dtype, const uint8_t* mask, const void* q, const void* k,
                         const void* v, int64_t sb, int64_t sn, int64_t shh) {
  auto al = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
  return dtype == FAVIT_BF16 && mask == nullptr && hd == HD && window <= 15 && N >= 1 && N <= 400 && al(q) && al(k) &&
         al(v) && sb % 8 == 0 && sn % 8 == 0 && shh % 8 == 0 && encode_fn() != nullptr;
}

int attn_seq_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int window,
                 float scale, int64_t sb, int64_t sn, int64_t shh, cudaStream_t st) {
  SeqParams p = make_params(B, H, N, window, scale, sb, sn, shh, nullptr, 3);
  CUtensorMap tq, tk, tv;
  if (int rc = make_map(&tq, q, B, H, N, sb, sn, shh, p.rows_box)) return rc;
  if (int rc = make_map(&tk, k, B, H, N, sb, sn, shh, p.rows_box)) return rc;
  if (int rc = make_map(&tv, v, B, H, N, sb, sn, shh, p.rows_box)) return rc;
  const size_t smem = (size_t)p.G * 3 * p.alloc_rows * kRowBytes + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_seq_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  const unsigned grid = (unsigned)ceil_div64(p.pairs, p.G);
  p.ahead = prefetch_distance((const void*)attn_seq_fwd_kernel<4>, p.G * p.sh.tiles * 32, smem, grid, 50);
  attn_seq_fwd_kernel<4><<<grid, p.G * p.sh.tiles * 32, smem, st>>>(tq, tk, tv, (__nv_bfloat16*)out, lse, p);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

int attn_seq_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout, void* dq,
                 void* dk, void* dv, float* colsum, int B, int H, int N, int window, float scale, int64_t sb, int64_t sn,
                 int64_t shh, cudaStream_t st) {
  SeqParams p = make_params(B, H, N, window, scale, sb, sn, shh, colsum, 4);
  CUtensorMap tq, tk, tv, td;
  if (int rc = make_map(&tq, q, B, H, N, sb, sn, shh, p.rows_box)) return rc;
  if (int rc = make_map(&tk, k, B, H, N, sb, sn, shh, p.rows_box)) return rc;
  if (int rc = make_map(&tv, v, B, H, N, sb, sn, shh, p.rows_box)) return rc;
  if (int rc = make_map(&td, dout, B, H, N, (int64_t)N * H * HD, (int64_t)H * HD, HD, p.rows_box)) return rc;
  const size_t smem = (size_t)p.G * (4 * p.alloc_rows * kRowBytes + 2 * p.alloc_rows * 4 + 3 * HD * 4) + 32 + 1024;
  static bool configured = false;
  if (!configured) {
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_seq_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = true;
  }
  const unsigned grid = (unsigned)ceil_div64(p.pairs, p.G);
  const unsigned threads = p.G * p.sh.tiles * 32;
  p.ahead = prefetch_distance((const void*)attn_seq_bwd_kernel<4>, (int)threads, smem, grid, 0);
  attn_seq_bwd_kernel<4><<<grid, threads, smem, st>>>(tq, tk, tv, td, (const __nv_bfloat16*)o, lse, (__nv_bfloat16*)dq,
                                                      (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, p);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

}  // namespace favit
