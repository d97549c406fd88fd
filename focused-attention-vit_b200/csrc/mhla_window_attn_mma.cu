// MHLA windowed attention core on tensor-core tiles (bf16, no mask, window <= 15, head_dim 32/64/128), sm_100a.
//
// Same math and the same reference span as mhla_window_attn.cu (/root/reference/models/mhla.py:109-154: banded softmax
// with the duplicated-edge multiplicities of mhla.py:72-79), reorganised so that the contractions are 16 x 8 x 16
// tensor-core tiles instead of per-lane FMAs + shuffles:
//   one warp owns 16 consecutive queries (or, in the dK/dV pass, 16 consecutive keys) of one (image, head);
//   the <= 16 + W - 1 band rows it needs, plus up to two "edge" rows (key N-1 / key 0, which early / late queries
//   attend several times), are staged once in shared memory with 16-byte cp.async copies (each K/V/Q/dO row is read
//   from HBM/L2 exactly once per tile instead of once per query), fragments come from ldmatrix, S = Q.K^T and P.V (and
//   the four backward contractions) run as mma.sync.m16n8k16 with fp32 accumulation, and the softmax works on the
//   accumulator fragments with two quad shuffles per row.  Outputs are staged through shared memory and written as
//   full 16-byte row segments.
// A band of 7 keys is far too narrow for a 128-row tcgen05 tile (the dense 128 x 134 score block would be 95 % zeros);
// the op stays HBM-bound and the point of the tensor-core formulation is only to get the instruction
// stream out of the way of the memory pipeline.  Masked, fp32 and wide-window calls use the SIMT kernels.
#include "attn_mma_common.cuh"

namespace favit {
namespace {

using namespace attn;

constexpr int kWarps = 4;

// =================================================================================================================
// forward
// =================================================================================================================
template <int HD, int NT>
__global__ void __launch_bounds__(kWarps * 32) attn_mma_fwd_kernel(const __nv_bfloat16* __restrict__ q,
                                                                  const __nv_bfloat16* __restrict__ k,
                                                                  const __nv_bfloat16* __restrict__ v,
                                                                  __nv_bfloat16* __restrict__ out,
                                                                  float* __restrict__ lse, Shape sh) {
  constexpr int NK = NT * 8;
  constexpr int kWarpBytes = (16 + 2 * NK) * HD * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wt = (int64_t)blockIdx.x * kWarps + warp;
  if (wt >= (int64_t)sh.B * sh.H * sh.tiles) return;
  uint8_t* sQ = smem + warp * kWarpBytes;
  uint8_t* sK = sQ + 16 * HD * 2;
  uint8_t* sV = sK + NK * HD * 2;
  const int qt = (int)(wt % sh.tiles);
  const int h = (int)((wt / sh.tiles) % sh.H);
  const int b = (int)(wt / ((int64_t)sh.tiles * sh.H));
  const int N = sh.N, i0 = qt * 16;
  const int64_t base = (int64_t)b * sh.sb + (int64_t)h * sh.sh;
  const KeySlots ks = key_slots(i0, N, sh.W);

  stage_rows<HD>(sQ, 16, lane, q, [&](int r) { return (i0 + r < N) ? q + base + (int64_t)(i0 + r) * sh.sn : nullptr; });
  stage_rows<HD>(sK, NK, lane, q, [&](int s) { const int j = ks.key(s); return j >= 0 ? k + base + (int64_t)j * sh.sn : nullptr; });
  stage_rows<HD>(sV, NK, lane, q, [&](int s) { const int j = ks.key(s); return j >= 0 ? v + base + (int64_t)j * sh.sn : nullptr; });
  cp_async_wait_all();
  __syncwarp();

  float s[NT][4];
  scores<HD, NT>(sQ, sK, lane, s);

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
      const float p = exp2f(s[nt][e] - (e < 2 ? mx0 : mx1));
      s[nt][e] = p;
      if (e < 2) sum0 += p; else sum1 += p;
    }
  sum0 = quad_sum(sum0);
  sum1 = quad_sum(sum1);

  float o[HD / 8][4];
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nd][e] = 0.f;
  pv<HD, NT>(s, sV, lane, o);

  __syncwarp();  // every lane is done reading sQ
  stage_acc<HD>(sQ, o, 1.f / sum0, 1.f / sum1, lane);
  __syncwarp();
  store_rows<HD>(sQ, lane, [&](int r) {
    return (i0 + r < N) ? out + (((int64_t)b * N + i0 + r) * sh.H + h) * HD : nullptr;
  });
  if ((lane & 3) == 0) {
    float* l = lse + ((int64_t)b * sh.H + h) * N;
    if (i0 + r0 < N) l[i0 + r0] = (mx0 + log2f(sum0)) * kLn2;
    if (i0 + r0 + 8 < N) l[i0 + r0 + 8] = (mx1 + log2f(sum1)) * kLn2;
  }
}

// =================================================================================================================
// backward, query-major pass: dQ and delta_i = dO_i . O_i
// =================================================================================================================
template <int HD, int NT>
__global__ void __launch_bounds__(kWarps * 32) attn_mma_dq_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    const __nv_bfloat16* __restrict__ o, const float* __restrict__ lse, const __nv_bfloat16* __restrict__ dout,
    __nv_bfloat16* __restrict__ dq, float* __restrict__ delta, Shape sh) {
  constexpr int NK = NT * 8;
  constexpr int kWarpBytes = (3 * 16 + 2 * NK) * HD * 2;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wt = (int64_t)blockIdx.x * kWarps + warp;
  if (wt >= (int64_t)sh.B * sh.H * sh.tiles) return;
  uint8_t* sQ = smem + warp * kWarpBytes;
  uint8_t* sdO = sQ + 16 * HD * 2;
  uint8_t* sO = sdO + 16 * HD * 2;   // O tile, later the dQ staging tile
  uint8_t* sK = sO + 16 * HD * 2;
  uint8_t* sV = sK + NK * HD * 2;
  const int qt = (int)(wt % sh.tiles);
  const int h = (int)((wt / sh.tiles) % sh.H);
  const int b = (int)(wt / ((int64_t)sh.tiles * sh.H));
  const int N = sh.N, i0 = qt * 16;
  const int64_t base = (int64_t)b * sh.sb + (int64_t)h * sh.sh;
  const KeySlots ks = key_slots(i0, N, sh.W);
  auto orow = [&](int r) { return (((int64_t)b * N + i0 + r) * sh.H + h) * HD; };

  stage_rows<HD>(sQ, 16, lane, q, [&](int r) { return (i0 + r < N) ? q + base + (int64_t)(i0 + r) * sh.sn : nullptr; });
  stage_rows<HD>(sdO, 16, lane, q, [&](int r) { return (i0 + r < N) ? dout + orow(r) : nullptr; });
  stage_rows<HD>(sO, 16, lane, q, [&](int r) { return (i0 + r < N) ? o + orow(r) : nullptr; });
  stage_rows<HD>(sK, NK, lane, q, [&](int s) { const int j = ks.key(s); return j >= 0 ? k + base + (int64_t)j * sh.sn : nullptr; });
  stage_rows<HD>(sV, NK, lane, q, [&](int s) { const int j = ks.key(s); return j >= 0 ? v + base + (int64_t)j * sh.sn : nullptr; });
  cp_async_wait_all();
  __syncwarp();

  // delta: the A fragments of dO and O hold the same (row, column) elements
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t a[4], c[4];
    load_a<HD>(sdO, kk, lane, a);
    load_a<HD>(sO, kk, lane, c);
    d0 += bf16x2_dot(a[0], c[0]) + bf16x2_dot(a[2], c[2]);
    d1 += bf16x2_dot(a[1], c[1]) + bf16x2_dot(a[3], c[3]);
  }
  d0 = quad_sum(d0);
  d1 = quad_sum(d1);

  float s[NT][4], dp[NT][4];
  scores<HD, NT>(sQ, sK, lane, s);
  scores<HD, NT>(sdO, sV, lane, dp);

  const int r0 = lane >> 2;
  const bool ok0 = i0 + r0 < N, ok1 = i0 + r0 + 8 < N;
  const RowSlots w0 = row_slots(ks, min(i0 + r0, N - 1), N, sh.W), w1 = row_slots(ks, min(i0 + r0 + 8, N - 1), N, sh.W);
  const float* l = lse + ((int64_t)b * sh.H + h) * N;
  // rows past the end of the sequence get L = +inf, i.e. P = 0
  const float L0 = ok0 ? l[i0 + r0] * kLog2e : CUDART_INF_F, L1 = ok1 ? l[i0 + r0 + 8] * kLog2e : CUDART_INF_F;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int slot = nt * 8 + (lane & 3) * 2 + (e & 1);
      const float p = exp2f(fmaf(s[nt][e], sh.scale_log2, (e < 2 ? w0 : w1).bias(slot)) - (e < 2 ? L0 : L1));
      s[nt][e] = p * (dp[nt][e] - (e < 2 ? d0 : d1));
    }
  float acc[HD / 8][4];
#pragma unroll
  for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
  pv<HD, NT>(s, sK, lane, acc);

  __syncwarp();
  stage_acc<HD>(sO, acc, sh.scale, sh.scale, lane);
  __syncwarp();
  store_rows<HD>(sO, lane, [&](int r) { return (i0 + r < N) ? dq + base + (int64_t)(i0 + r) * sh.sn : nullptr; });
  if (sh.colsum) tile_colsum<HD>(sO, lane, sh.colsum + h * HD);
  if ((lane & 3) == 0) {
    float* dl = delta + ((int64_t)b * sh.H + h) * N;
    if (ok0) dl[i0 + r0] = d0;
    if (ok1) dl[i0 + r0 + 8] = d1;
  }
}

// =================================================================================================================
// backward, key-major pass: dK and dV gathered from every query whose window holds the key (no atomics)
// =================================================================================================================
template <int HD, int NT>
__global__ void __launch_bounds__(kWarps * 32) attn_mma_dkv_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
    const float* __restrict__ lse, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ delta,
    __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, Shape sh) {
  constexpr int NQ = NT * 8;
  constexpr int kWarpBytes = (2 * 16 + 2 * NQ) * HD * 2 + 2 * NQ * 4;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wt = (int64_t)blockIdx.x * kWarps + warp;
  if (wt >= (int64_t)sh.B * sh.H * sh.tiles) return;
  uint8_t* sK = smem + warp * kWarpBytes;   // K tile, later the dK staging tile
  uint8_t* sV = sK + 16 * HD * 2;           // V tile, later the dV staging tile
  uint8_t* sQ = sV + 16 * HD * 2;
  uint8_t* sdO = sQ + NQ * HD * 2;
  float* sL = reinterpret_cast<float*>(sdO + NQ * HD * 2);
  float* sD = sL + NQ;
  const int kt = (int)(wt % sh.tiles);
  const int h = (int)((wt / sh.tiles) % sh.H);
  const int b = (int)(wt / ((int64_t)sh.tiles * sh.H));
  const int N = sh.N, j0 = kt * 16;
  const int64_t base = (int64_t)b * sh.sb + (int64_t)h * sh.sh;
  const QuerySlots qs = query_slots(j0, N, sh.W);

  stage_rows<HD>(sK, 16, lane, q, [&](int r) { return (j0 + r < N) ? k + base + (int64_t)(j0 + r) * sh.sn : nullptr; });
  stage_rows<HD>(sV, 16, lane, q, [&](int r) { return (j0 + r < N) ? v + base + (int64_t)(j0 + r) * sh.sn : nullptr; });
  stage_rows<HD>(sQ, NQ, lane, q, [&](int s) { const int i = qs.query(s); return i >= 0 ? q + base + (int64_t)i * sh.sn : nullptr; });
  stage_rows<HD>(sdO, NQ, lane, q, [&](int s) {
    const int i = qs.query(s);
    return i >= 0 ? dout + (((int64_t)b * N + i) * sh.H + h) * HD : nullptr;
  });
  for (int s = lane; s < NQ; s += 32) {
    const int i = qs.query(s);
    sL[s] = i >= 0 ? lse[((int64_t)b * sh.H + h) * N + i] * kLog2e : CUDART_INF_F;
    sD[s] = i >= 0 ? delta[((int64_t)b * sh.H + h) * N + i] : 0.f;
  }
  cp_async_wait_all();
  __syncwarp();

  float s[NT][4], dp[NT][4];
  scores<HD, NT>(sK, sQ, lane, s);    // S^T[key][query]
  scores<HD, NT>(sV, sdO, lane, dp);  // dP^T[key][query]

  const int r0 = lane >> 2;
  const int ja = j0 + r0, jb = j0 + r0 + 8;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int slot = nt * 8 + (lane & 3) * 2 + c;
      const int i = qs.query(slot);
      const WindowRow w = window_row(max(i, 0), N, sh.W);
      // unused slots carry L = +inf (P = 0); keys past the end of the sequence are never stored
      const float L = sL[slot], dl = sD[slot];
      const float lb_in = kLog2Int[1 + w.pad], lb_out = kLog2Int[w.pad];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int e = half * 2 + c;
        const int j = half ? jb : ja;
        const bool in_band = (unsigned)(j - w.s) < (unsigned)(w.e - w.s);
        const bool edge = j == w.tgt;
        const float bias = in_band ? (edge ? lb_in : 0.f) : (edge ? lb_out : -CUDART_INF_F);
        const float p = exp2f(fmaf(s[nt][e], sh.scale_log2, bias) - L);
        s[nt][e] = p;
        dp[nt][e] = p * (dp[nt][e] - dl);
      }
    }
  __syncwarp();  // K / V tiles are dead from here on: reuse them as staging
  {
    float acc[HD / 8][4];
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
    pv<HD, NT>(s, sdO, lane, acc);  // dV = P^T . dO
    stage_acc<HD>(sV, acc, 1.f, 1.f, lane);
  }
  {
    float acc[HD / 8][4];
#pragma unroll
    for (int nd = 0; nd < HD / 8; ++nd)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nd][e] = 0.f;
    pv<HD, NT>(dp, sQ, lane, acc);  // dK = scale . dS^T . Q
    stage_acc<HD>(sK, acc, sh.scale, sh.scale, lane);
  }
  __syncwarp();
  store_rows<HD>(sV, lane, [&](int r) { return (j0 + r < N) ? dv + base + (int64_t)(j0 + r) * sh.sn : nullptr; });
  store_rows<HD>(sK, lane, [&](int r) { return (j0 + r < N) ? dk + base + (int64_t)(j0 + r) * sh.sn : nullptr; });
  if (sh.colsum) {
    tile_colsum<HD>(sK, lane, sh.colsum + (sh.H + h) * HD);
    tile_colsum<HD>(sV, lane, sh.colsum + (2 * sh.H + h) * HD);
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return FAVIT_OK;
}

template <int HD>
int launch_fwd(const void* q, const void* k, const void* v, void* out, float* lse, const Shape& sh, cudaStream_t st) {
  constexpr int NT = 4;
  const size_t smem = (size_t)kWarps * (16 + 2 * NT * 8) * HD * 2;
  if (int rc = set_smem(attn_mma_fwd_kernel<HD, NT>, smem)) return rc;
  const int64_t wt = (int64_t)sh.B * sh.H * sh.tiles;
  attn_mma_fwd_kernel<HD, NT><<<(unsigned)ceil_div64(wt, kWarps), kWarps * 32, smem, st>>>(
      (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (__nv_bfloat16*)out, lse, sh);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

template <int HD, int NTK>
int launch_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout, void* dq,
               void* dk, void* dv, float* delta, const Shape& sh, cudaStream_t st) {
  constexpr int NT = 4;
  const int64_t wt = (int64_t)sh.B * sh.H * sh.tiles;
  const unsigned blocks = (unsigned)ceil_div64(wt, kWarps);
  {
    const size_t smem = (size_t)kWarps * (3 * 16 + 2 * NT * 8) * HD * 2;
    if (int rc = set_smem(attn_mma_dq_kernel<HD, NT>, smem)) return rc;
    attn_mma_dq_kernel<HD, NT><<<blocks, kWarps * 32, smem, st>>>(
        (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (const __nv_bfloat16*)o, lse,
        (const __nv_bfloat16*)dout, (__nv_bfloat16*)dq, delta, sh);
    FAVIT_CHECK_LAUNCH();
  }
  {
    const size_t smem = (size_t)kWarps * ((2 * 16 + 2 * NTK * 8) * HD * 2 + 2 * NTK * 8 * 4);
    if (int rc = set_smem(attn_mma_dkv_kernel<HD, NTK>, smem)) return rc;
    attn_mma_dkv_kernel<HD, NTK><<<blocks, kWarps * 32, smem, st>>>(
        (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, lse, (const __nv_bfloat16*)dout,
        delta, (__nv_bfloat16*)dk, (__nv_bfloat16*)dv, sh);
    FAVIT_CHECK_LAUNCH();
  }
  return FAVIT_OK;
}

}  // namespace

// Returns FAVIT_ERR_UNSUPPORTED (without setting an error) when the call must take the SIMT kernels instead.
bool attn_mma_applicable(int hd, int window, favit_dtype dtype, const uint8_t* mask) {
  return dtype == FAVIT_BF16 && mask == nullptr && (hd == 32 || hd == 64 || hd == 128) && window <= 15;
}

int attn_mma_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int hd,
                 int window, float scale, int64_t sb, int64_t sn, int64_t shh, cudaStream_t st) {
  Shape sh{B, H, N, window, sb, sn, shh, scale * kLog2e, scale, ceil_div(N, 16), nullptr};
  switch (hd) {
    case 32: return launch_fwd<32>(q, k, v, out, lse, sh, st);
    case 64: return launch_fwd<64>(q, k, v, out, lse, sh, st);
    default: return launch_fwd<128>(q, k, v, out, lse, sh, st);
  }
}

int attn_mma_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                 void* dq, void* dk, void* dv, float* delta, float* colsum, int B, int H, int N, int hd, int window,
                 float scale, int64_t sb, int64_t sn, int64_t shh, cudaStream_t st) {
  Shape sh{B, H, N, window, sb, sn, shh, scale * kLog2e, scale, ceil_div(N, 16), colsum};
  const bool wide = (17 + 4 * (window >> 1)) > 32;  // query slots a key tile may need
#define FAVIT_BWD(HD)                                                                                  \
  return wide ? launch_bwd<HD, 6>(q, k, v, o, lse, dout, dq, dk, dv, delta, sh, st)                    \
              : launch_bwd<HD, 4>(q, k, v, o, lse, dout, dq, dk, dv, delta, sh, st)
  switch (hd) {
    case 32: FAVIT_BWD(32);
    case 64: FAVIT_BWD(64);
    default: FAVIT_BWD(128);
  }
#undef FAVIT_BWD
}

}  // namespace favit
