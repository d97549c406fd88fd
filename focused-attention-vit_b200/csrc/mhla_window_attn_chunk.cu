// MHLA windowed attention in sequence CHUNKS (bf16, head_dim 64, no mask, window <= 15; forward N > 48, backward N > 400), sm_100a.
//
// Same math and reference span as mhla_window_attn_seq.cu (/root/reference/models/mhla.py:109-154, banded softmax with
// the duplicated-edge multiplicities of mhla.py:72-79) and the same TMA + mma.sync tile arithmetic; what changes is the
// unit of work.  A CTA owns a CHUNK of C = 16 T consecutive rows of one (image, head) sequence (forward C <= 64, backward
// C <= 176, chosen so that the chunks of a sequence are balanced) instead of whole sequences: long sequences do not fit
// in shared memory, and in the forward pass small CTAs hide the load latency better at every length:
//   forward : Q rows [c0, c0 + C), K / V rows [c0 - 8, c0 + C + 8) (the band never reaches further for W <= 15), one
//             TMA box per operand — rows before the first / past the last row of the sequence are zero-filled by the
//             tensor map — plus the two duplicated edge rows (key N-1, key 0), which one warp copies into two spare rows
//             of the K / V tiles.  One warp per 16-query tile, as in the whole-sequence kernel.
//   backward: ONE kernel per chunk as well.  The chunk owns the dQ rows and the dK / dV rows [c0, c0 + C).  dK / dV of a
//             key need P / dS of the queries within h of it, so phase A (scores, P, dS; dQ for the own tiles) also runs on
//             one HALO query tile on either side (T + 2 warps; Q / dO / O / LSE rows [c0 - 16, c0 + C + 16)): 2 / T of the
//             tile arithmetic and ~15 % of the reads are repeated by the neighbour chunk (out of L2: the neighbours run
//             at the same time), nothing is exchanged between CTAs and nothing is accumulated atomically.  The
//             duplicated edge keys receive from queries at the OTHER end of the sequence (key N-1 from queries < h, key
//             0 from queries > N-1-h): the tile that owns the edge key recomputes those <= h probabilities on the CUDA
//             cores (three 64-wide dot products per query, operands out of L2) and adds them to its accumulators before
//             they are rounded — bit-for-bit deterministic, no read-modify-write pass.
// The op is HBM-bound (AI = W/2 FLOP/B): algorithmic bytes fwd = 4*B*N*D*2, bwd = 8*B*N*D*2.
#include <cuda.h>

#include <algorithm>

#include "attn_seq_common.cuh"

namespace favit {
namespace {

using namespace attn;
using namespace seqk;

constexpr int kMaxChunkTiles = 11;   // backward: 13 warps, 4 x 208 rows x 128 B = 104 KB of tiles, two CTAs per SM
constexpr int kFwdChunkTiles = 4;    // forward: 64-row chunks, 33 KB of tiles, six CTAs per SM (sweep in profiles/r2_attn_seq_probe.txt:
                                     // 2 / 3 / 4 / 6 / 8 / 11 tiles -> 0.81 / 0.82 / 0.81 / 0.74 / 0.77 / 0.77 of HBM at ViT-B/16)

struct ChunkParams {
  Shape sh;
  int C;           // rows per chunk (multiple of 16)
  int T;           // 16-row tiles per chunk
  int nchunks;     // chunks per sequence
  int rows_kv;     // rows of a K / V box: C + 16
  int alloc_rows;  // rows of a K / V tile (forward) and of every tile (backward): C + 32
  int64_t pairs;   // B * H
  int cs_shared;   // backward: column sums through the CTA's shared-memory accumulator
  const __nv_bfloat16* q;     // raw pointers (strides in sh): the edge rows and the edge terms are read directly
  const __nv_bfloat16* k;
  const __nv_bfloat16* v;
};

// sequence key j -> row of the K / V tile whose row 0 is key kb.  Keys outside the loaded rows can only be the duplicated
// edge keys (spare rows E0 = key N-1, E1 = key 0) or slots whose probability is zero (any finite row will do: E0).
__device__ __forceinline__ int kv_row(int j, int kb, int rows_kv) {
  const int r = j - kb;
  if (j >= 0 && r >= 0 && r < rows_kv) return r;
  return rows_kv + (j == 0 ? 1 : 0);
}

// warp 0: K / V rows N-1 and 0 of this sequence -> spare rows rows_kv, rows_kv + 1 of the two tiles (one uint4 per lane)
__device__ __forceinline__ void load_edge_rows(uint8_t* sK, uint8_t* sV, const ChunkParams& p, int64_t base, int lane) {
  const int t = lane >> 4, e = (lane >> 3) & 1, c = lane & 7;
  const __nv_bfloat16* src = (t ? p.v : p.k) + base + (int64_t)(e ? 0 : p.sh.N - 1) * p.sh.sn + c * 8;
  *reinterpret_cast<uint4*>((t ? sV : sK) + row_off(p.rows_kv + e, c)) = *reinterpret_cast<const uint4*>(src);
}

// =================================================================================================================
// forward
// =================================================================================================================
template <int NT>
__global__ void __launch_bounds__(kMaxChunkTiles * 32, 2) attn_chunk_fwd_kernel(const __grid_constant__ CUtensorMap tmq,
                                                                            const __grid_constant__ CUtensorMap tmk,
                                                                            const __grid_constant__ CUtensorMap tmv,
                                                                            __nv_bfloat16* __restrict__ out,
                                                                            float* __restrict__ lse, ChunkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // the 128B swizzle is keyed on address bits
  const Shape& sh = p.sh;
  const size_t q_bytes = (size_t)p.C * kRowBytes, kv_bytes = (size_t)p.alloc_rows * kRowBytes;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + q_bytes;
  uint8_t* sV = sK + kv_bytes;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sV + kv_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pr = blockIdx.x / p.nchunks;
  const int c0 = (int)(blockIdx.x % p.nchunks) * p.C, kb = c0 - 8;
  const int b = (int)(pr / sh.H), h = (int)(pr % sh.H);
  const int64_t base = (int64_t)b * sh.sb + (int64_t)h * sh.sh;
  if (threadIdx.x == 0) {
    const uint32_t br = smem_u32(bar);
    ptx::mbar_init(br, 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
    ptx::mbar_expect_tx(br, (uint32_t)((p.C + 2 * p.rows_kv) * kRowBytes));
    tma_load_4d(smem_u32(sQ), &tmq, br, 0, h, c0, b);
    tma_load_4d(smem_u32(sK), &tmk, br, 0, h, kb, b);
    tma_load_4d(smem_u32(sV), &tmv, br, 0, h, kb, b);
  }
  if (warp == 0) load_edge_rows(sK, sV, p, base, lane);
  __syncthreads();
  ptx::mbar_wait(smem_u32(bar), 0);

  const int N = sh.N, i0 = c0 + warp * 16;
  if (i0 >= N) return;
  uint8_t* sQt = sQ + (size_t)warp * 16 * kRowBytes;  // own query tile, later the output staging tile
  const KeySlots ks = key_slots(i0, N, sh.W);
  int krow[NT / 2], vrow[NT / 2];
#pragma unroll
  for (int x = 0; x < NT / 2; ++x) {
    krow[x] = kv_row(ks.key(16 * x + (lane & 7) + 8 * (lane >> 4)), kb, p.rows_kv);
    vrow[x] = kv_row(ks.key(16 * x + (lane & 7) + 8 * ((lane >> 3) & 1)), kb, p.rows_kv);
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
// backward.  Warp w <-> query tile w - 1 of the chunk (w = 0 and w = T + 1: the halo tiles), rows c0 + 16 (w - 1) ...;
// Q / dO tile row r <-> sequence row c0 - 16 + r, K / V tile row r <-> key c0 - 8 + r.  Phase A is the whole-sequence
// kernel's (aligned key slots, edge slots 0 / 31 on the spare rows), its P / dS blocks are parked over the K tile
// ((T + 2) x 1 KB each: exactly the tile), phase B takes the blocks of warps w - 1, w, w + 1.
// =================================================================================================================
template <int NT>
__global__ void __launch_bounds__((kMaxChunkTiles + 2) * 32, 2) attn_chunk_bwd_kernel(
    const __grid_constant__ CUtensorMap tmq, const __grid_constant__ CUtensorMap tmk,
    const __grid_constant__ CUtensorMap tmv, const __grid_constant__ CUtensorMap tmdo,
    const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
    __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, ChunkParams p) {
  static_assert(NT == 4, "32 aligned key slots per query tile");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const Shape& sh = p.sh;
  const size_t tile_bytes = (size_t)p.alloc_rows * kRowBytes;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + tile_bytes;    // after phase A: the parked P blocks [T + 2][16][32], then the dS blocks
  uint8_t* sV = sK + tile_bytes;    // after phase A: output staging
  uint8_t* sdO = sV + tile_bytes;
  float* sL = reinterpret_cast<float*>(sdO + tile_bytes);  // [alloc_rows] log2-domain LSE (+inf outside the sequence)
  float* sD = sL + p.alloc_rows;                           // [alloc_rows] delta
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sD + p.alloc_rows);
  uint8_t* zero16 = reinterpret_cast<uint8_t*>(bar + 2);   // 16 zero bytes: the absent 8 x 8 blocks
  float* sCS = reinterpret_cast<float*>(zero16 + 16);      // [3 * 64] column sums of dQ | dK | dV of this chunk
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pr = blockIdx.x / p.nchunks;
  const int chunk = (int)(blockIdx.x % p.nchunks);
  const int c0 = chunk * p.C, qb = c0 - 16, kb = c0 - 8;
  const int b = (int)(pr / sh.H), h = (int)(pr % sh.H);
  const int64_t base = (int64_t)b * sh.sb + (int64_t)h * sh.sh;
  const int N = sh.N;
  if (threadIdx.x == 0) {
    const uint32_t br = smem_u32(bar);
    ptx::mbar_init(br, 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
    ptx::mbar_expect_tx(br, (uint32_t)((2 * p.alloc_rows + 2 * p.rows_kv) * kRowBytes));
    tma_load_4d(smem_u32(sQ), &tmq, br, 0, h, qb, b);
    tma_load_4d(smem_u32(sdO), &tmdo, br, 0, h, qb, b);
    tma_load_4d(smem_u32(sK), &tmk, br, 0, h, kb, b);
    tma_load_4d(smem_u32(sV), &tmv, br, 0, h, kb, b);
  }
  if (warp == 0) load_edge_rows(sK, sV, p, base, lane);
  if (threadIdx.x < 4) reinterpret_cast<uint32_t*>(zero16)[threadIdx.x] = 0u;
  for (int i = threadIdx.x; i < 3 * HD; i += blockDim.x) sCS[i] = 0.f;

  const int i0 = c0 + (warp - 1) * 16;         // first query (phase A) / key (phase B) of this warp's tile
  const int lr0 = warp * 16;                   // its first row in the Q / dO tiles and in sL / sD
  const bool live = i0 >= 0 && i0 < N;         // a tile outside the sequence only parks zeros
  const bool own = live && warp >= 1 && warp <= p.T;
  // O (for delta) and the LSE come straight from global: fetch them while the TMA boxes are in flight
  uint4 o_pre[4];
  float lse_pre = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = (lane >> 3) + 4 * j, c = lane & 7;
    o_pre[j] = make_uint4(0, 0, 0, 0);
    if (live && i0 + r < N) o_pre[j] = *reinterpret_cast<const uint4*>(o + (((int64_t)b * N + i0 + r) * sh.H + h) * HD + c * 8);
  }
  if (live && lane < 16 && i0 + lane < N) lse_pre = lse[((int64_t)b * sh.H + h) * N + i0 + lane];
  __syncthreads();
  ptx::mbar_wait(smem_u32(bar), 0);

  const int r0 = lane >> 2;
  uint32_t p_pk[NT][2], ds_pk[NT][2];  // this tile's P and dS, bf16 pairs in accumulator-fragment order
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) p_pk[nt][0] = p_pk[nt][1] = ds_pk[nt][0] = ds_pk[nt][1] = 0u;

  // ---------------------------------------------------------------- phase A: queries i0 .. i0+15
  if (live) {
    const uint8_t* sQt = sQ + (size_t)lr0 * kRowBytes;
    const uint8_t* sdOt = sdO + (size_t)lr0 * kRowBytes;
    {
      if (lane < 16) sL[lr0 + lane] = (i0 + lane < N) ? lse_pre * kLog2e : CUDART_INF_F;  // +inf -> P = 0
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
        if (c == 0) sD[lr0 + r] = part;
      }
      __syncwarp();
    }
    const AlSlots ks = al_slots(i0, N, sh.W);
    int krow[NT / 2], vrow[NT / 2];
#pragma unroll
    for (int x = 0; x < NT / 2; ++x) {
      krow[x] = kv_row(ks.key(16 * x + (lane & 7) + 8 * (lane >> 4)), kb, p.rows_kv);
      vrow[x] = kv_row(ks.key(16 * x + (lane & 7) + 8 * ((lane >> 3) & 1)), kb, p.rows_kv);
    }
    float s[NT][4], dp[NT][4];
    scores_rows<NT>(sQt, sK, krow, lane, s);
    scores_rows<NT>(sdOt, sV, krow, lane, dp);
    const bool ok0 = i0 + r0 < N, ok1 = i0 + r0 + 8 < N;
    const RowSlots w0 = al_row_slots(ks, min(i0 + r0, N - 1), N, sh.W);
    const RowSlots w1 = al_row_slots(ks, min(i0 + r0 + 8, N - 1), N, sh.W);
    const float L0 = sL[lr0 + r0], L1 = sL[lr0 + r0 + 8], d0 = sD[lr0 + r0], d1 = sD[lr0 + r0 + 8];
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
    if (own) {  // dQ = scale . dS . K (the halo tiles' dQ belongs to the neighbour chunk)
      float acc[HD / 8][4];
#pragma unroll
      for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < NT / 2; ++kk) {
        const uint32_t a[4] = {ds_pk[2 * kk][0], ds_pk[2 * kk][1], ds_pk[2 * kk + 1][0], ds_pk[2 * kk + 1][1]};
        mma_rows(a, sK, vrow[kk], lane, acc);
      }
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
      if (sh.colsum) colsum_frags(v0, v1, lane, p.cs_shared ? sCS : sh.colsum + h * HD, p.cs_shared);
    }
  }
  __syncthreads();  // nobody reads K / V rows any more
  {                 // park P in the K tile's first half, dS in its second half (dead tiles park zeros)
    uint8_t* pP = sK + (size_t)warp * kBlkBytes;
    uint8_t* pS = pP + (size_t)(p.T + 2) * kBlkBytes;
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

  // ---------------------------------------------------------------- phase B: keys i0 .. i0+15 of the own tiles
  if (own) {
    const int j0 = i0;
    const uint8_t* sP = sK;
    const uint8_t* sS = sK + (size_t)(p.T + 2) * kBlkBytes;
    const uint32_t zaddr = smem_u32(zero16);
    // ldmatrix.x4.trans row addresses of the A fragments P^T / dS^T (offsets relative to sP / sS), blocks of warps w-1, w, w+1:
    //   k-step 0 (queries [j0-8, j0+8)):  m0 (w-1, rows 8.., chunk 3)  m1 zero                   m2 (w, rows 0.., chunk 1)  m3 (w, rows 0.., chunk 2)
    //   k-step 1 (queries [j0+8, j0+24)): m0 (w, rows 8.., chunk 1)    m1 (w, rows 8.., chunk 2)  m2 zero                    m3 (w+1, rows 0.., chunk 0)
    const int mi = lane >> 3, r8 = lane & 7;
    int offA[2];
    {
      const int q0 = (mi == 0) ? warp - 1 : warp, rr0 = (mi == 0) ? 8 + r8 : r8, cc0 = (mi == 0) ? 3 : mi - 1;
      offA[0] = (mi == 1) ? -1 : q0 * kBlkBytes + (int)tile_off<32>(rr0, cc0);
      const int q1 = (mi == 3) ? warp + 1 : warp, rr1 = (mi == 3) ? r8 : 8 + r8, cc1 = (mi == 3) ? 0 : mi + 1;
      offA[1] = (mi == 2) ? -1 : q1 * kBlkBytes + (int)tile_off<32>(rr1, cc1);
    }
    int brow[2];   // Q / dO tile rows of the queries of a k-step: sequence row j0 - 8 + 16 kk + ... minus qb
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) brow[kk] = lr0 - 8 + 16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1);
    const int ja = j0 + r0, jb = j0 + r0 + 8;
    const bool okA = ja < N, okB = jb < N;

    // The duplicated edge key of this tile, if it owns one: key N-1 receives from queries t < h, key 0 from queries
    // N-1-t, with multiplicity pad = h - t.  pe = pad . exp(s - L), ds = pe (dP - delta), all recomputed here.
    const int hh = sh.W >> 1;
    const bool edgeA = j0 <= N - 1 && N - 1 < j0 + 16, edgeB = j0 == 0;
    const bool edge = edgeA || edgeB;               // warp-uniform (never both: N > 16)
    const int erow = edgeA ? N - 1 - j0 : 0;        // row of the edge key in this tile
    float e_pe[7], e_ds[7];
#pragma unroll
    for (int t = 0; t < 7; ++t) e_pe[t] = e_ds[t] = 0.f;
    if (edge) {
      auto ld2 = [&](const __nv_bfloat16* row, float& x, float& y) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(row + 2 * lane);
        x = __uint_as_float(u << 16);
        y = __uint_as_float(u & 0xffff0000u);
      };
      const int ekey = edgeA ? N - 1 : 0;
      float k0, k1, v0, v1;
      ld2(p.k + base + (int64_t)ekey * sh.sn, k0, k1);
      ld2(p.v + base + (int64_t)ekey * sh.sn, v0, v1);
#pragma unroll
      for (int t = 0; t < 7; ++t) {
        if (t < hh) {
          const int i = edgeA ? t : N - 1 - t;
          const int64_t orow = (((int64_t)b * N + i) * sh.H + h) * HD;
          float q0, q1, g0, g1, o0, o1;
          ld2(p.q + base + (int64_t)i * sh.sn, q0, q1);
          ld2(dout + orow, g0, g1);
          ld2(o + orow, o0, o1);
          float sd = q0 * k0 + q1 * k1, dpe = g0 * v0 + g1 * v1, dl = g0 * o0 + g1 * o1;
#pragma unroll
          for (int x = 16; x > 0; x >>= 1) {
            sd += __shfl_xor_sync(0xffffffffu, sd, x);
            dpe += __shfl_xor_sync(0xffffffffu, dpe, x);
            dl += __shfl_xor_sync(0xffffffffu, dl, x);
          }
          const float pe = (float)(hh - t) * exp2f(sd * sh.scale_log2 - lse[((int64_t)b * sh.H + h) * N + i] * kLog2e);
          e_pe[t] = pe;
          e_ds[t] = pe * (dpe - dl);
        }
      }
    }

    uint8_t* stage = sV + (size_t)(warp - 1) * 16 * kRowBytes;  // the V tile is free since the barrier
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
      if (edge && r0 == (erow & 7)) {  // the four lanes that hold row `erow` of the accumulators
        const int col = (lane & 3) * 2;
#pragma unroll
        for (int t = 0; t < 7; ++t) {
          if (t < hh) {
            const int i = edgeA ? t : N - 1 - t;
            const __nv_bfloat16* src = pass == 0 ? dout + (((int64_t)b * N + i) * sh.H + h) * HD : p.q + base + (int64_t)i * sh.sn;
            const float coef = pass == 0 ? e_pe[t] : e_ds[t];
#pragma unroll
            for (int nd = 0; nd < HD / 8; ++nd) {
              const uint32_t u = *reinterpret_cast<const uint32_t*>(src + nd * 8 + col);
              const float x = __uint_as_float(u << 16), y = __uint_as_float(u & 0xffff0000u);
              if (erow < 8) {
                acc[nd][0] = fmaf(coef, x, acc[nd][0]);
                acc[nd][1] = fmaf(coef, y, acc[nd][1]);
              } else {
                acc[nd][2] = fmaf(coef, x, acc[nd][2]);
                acc[nd][3] = fmaf(coef, y, acc[nd][3]);
              }
            }
          }
        }
      }
      const float m = pass == 0 ? 1.f : sh.scale;
      __syncwarp();
      stage_acc<HD>(stage, acc, okA ? m : 0.f, okB ? m : 0.f, lane);  // rows past N must stay exact zeros (column sums)
      __syncwarp();
      __nv_bfloat16* dst = pass == 0 ? dv : dk;
      store_rows<HD>(stage, lane, [&](int r) { return (j0 + r < N) ? dst + base + (int64_t)(j0 + r) * sh.sn : nullptr; });
      if (sh.colsum)
        colsum_staged(stage, lane, p.cs_shared ? sCS + (pass == 0 ? 2 : 1) * HD
                                               : sh.colsum + ((pass == 0 ? 2 : 1) * sh.H + h) * HD, p.cs_shared);
    }
  }
  if (sh.colsum && p.cs_shared) {  // one global atomic per column per CTA
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * HD; i += blockDim.x)
      atomicAdd(sh.colsum + ((i / HD) * sh.H + h) * HD + (i % HD), sCS[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
ChunkParams make_params(const void* q, const void* k, const void* v, int B, int H, int N, int window, float scale,
                        int64_t sb, int64_t sn, int64_t shh, float* colsum, int max_tiles) {
  ChunkParams p;
  const int cmax = max_tiles * 16;
  p.nchunks = ceil_div(N, cmax);
  p.T = ceil_div(ceil_div(N, p.nchunks), 16);   // balanced chunks
  p.C = p.T * 16;
  p.nchunks = ceil_div(N, p.C);
  p.sh = Shape{B, H, N, window, sb, sn, shh, scale * kLog2e, scale, ceil_div(N, 16), colsum};
  p.rows_kv = p.C + 16;
  p.alloc_rows = p.C + 32;
  p.pairs = (int64_t)B * H;
  p.cs_shared = H <= 8 ? 1 : 0;   // same crossover as the whole-sequence kernel
  p.q = (const __nv_bfloat16*)q;
  p.k = (const __nv_bfloat16*)k;
  p.v = (const __nv_bfloat16*)v;
  return p;
}

}  // namespace

// Shortest sequence that goes to the chunk kernels.  Forward: N > 48 — many small CTAs (64-row chunks, six per SM) overlap
// loads and arithmetic better than one CTA per sequence: 0.82 vs 0.73 of HBM at ViT-B/16 (N 197), 0.75 vs 0.73 at N 65; at
// N <= 48 the whole-sequence kernel packs several sequences per CTA.  Backward: past the whole-sequence kernel's limit
// (N > 400): the halo tiles make small chunks expensive (0.43 vs 0.53 at N 197), from N ~ 400 on the two are within
// +-7 % of each other.  FAVIT_CHUNK_MIN_N / FAVIT_CHUNK_FWD_TILES / FAVIT_CHUNK_BWD_TILES override (tuning only).
bool attn_chunk_applicable(int hd, int window, int N, int B, int H, favit_dtype dtype, const uint8_t* mask, const void* q,
                           const void* k, const void* v, int64_t sb, int64_t sn, int64_t shh, bool backward) {
  static const int forced = [] { const char* e = getenv("FAVIT_CHUNK_MIN_N"); return e ? std::max(atoi(e), 32) : 0; }();
  const int min_n = forced ? forced : (backward ? 400 : 48);
  auto al = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
  return dtype == FAVIT_BF16 && mask == nullptr && hd == HD && window >= 1 && window <= 15 && N > min_n &&
         (int64_t)B * H * ceil_div(N, 16) < INT32_MAX && al(q) && al(k) && al(v) && sb % 8 == 0 && sn % 8 == 0 &&
         shh % 8 == 0 && encode_fn() != nullptr;
}

int attn_chunk_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int window,
                   float scale, int64_t sb, int64_t sn, int64_t shh, cudaStream_t st) {
  static const int fwd_tiles = [] { const char* e = getenv("FAVIT_CHUNK_FWD_TILES"); return e ? std::min(std::max(atoi(e), 1), kMaxChunkTiles) : kFwdChunkTiles; }();
  const ChunkParams p = make_params(q, k, v, B, H, N, window, scale, sb, sn, shh, nullptr, fwd_tiles);
  CUtensorMap tq, tk, tv;
  if (int rc = make_map(&tq, q, B, H, N, sb, sn, shh, p.C)) return rc;
  if (int rc = make_map(&tk, k, B, H, N, sb, sn, shh, p.rows_kv)) return rc;
  if (int rc = make_map(&tv, v, B, H, N, sb, sn, shh, p.rows_kv)) return rc;
  const size_t smem = (size_t)(p.C + 2 * p.alloc_rows) * kRowBytes + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_chunk_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    configured = true;
  }
  const unsigned grid = (unsigned)(p.pairs * p.nchunks);
  attn_chunk_fwd_kernel<4><<<grid, p.T * 32, smem, st>>>(tq, tk, tv, (__nv_bfloat16*)out, lse, p);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

int attn_chunk_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout, void* dq,
                   void* dk, void* dv, float* colsum, int B, int H, int N, int window, float scale, int64_t sb, int64_t sn,
                   int64_t shh, cudaStream_t st) {
  static const int bwd_tiles = [] { const char* e = getenv("FAVIT_CHUNK_BWD_TILES"); return e ? std::min(std::max(atoi(e), 1), kMaxChunkTiles) : kMaxChunkTiles; }();
  const ChunkParams p = make_params(q, k, v, B, H, N, window, scale, sb, sn, shh, colsum, bwd_tiles);
  CUtensorMap tq, tk, tv, td;
  if (int rc = make_map(&tq, q, B, H, N, sb, sn, shh, p.alloc_rows)) return rc;
  if (int rc = make_map(&tk, k, B, H, N, sb, sn, shh, p.rows_kv)) return rc;
  if (int rc = make_map(&tv, v, B, H, N, sb, sn, shh, p.rows_kv)) return rc;
  if (int rc = make_map(&td, dout, B, H, N, (int64_t)N * H * HD, (int64_t)H * HD, HD, p.alloc_rows)) return rc;
  const size_t smem = (size_t)4 * p.alloc_rows * kRowBytes + 2 * (size_t)p.alloc_rows * 4 + 16 + 16 + 3 * HD * 4 + 1024;
  static bool configured = false;
  if (!configured) {
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(attn_chunk_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    configured = true;
  }
  const unsigned grid = (unsigned)(p.pairs * p.nchunks);
  attn_chunk_bwd_kernel<4><<<grid, (p.T + 2) * 32, smem, st>>>(tq, tk, tv, td, (const __nv_bfloat16*)o,
                                                               (const __nv_bfloat16*)dout, lse, (__nv_bfloat16*)dq,
                                                               (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, p);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

}  // namespace favit
