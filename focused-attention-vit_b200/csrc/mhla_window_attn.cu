// MHLA windowed attention core, forward + backward, for sm_100a.
//
// Replaces /root/reference/models/mhla.py:109-154 (window index table, K/V gathers, scaled scores, mask,
// softmax, PV) with a banded softmax carrying integer multiplicities: query i attends keys
// [max(0,i-h), min(N,i+h+1)) once each, plus `pad = W - len` extra copies of key N-1 (when the band is
// clipped on the left / the sequence is short) or of key 0 (when clipped on the right) — exactly the
// duplicated indices of mhla.py:72-79.  Nothing of shape [B,H,N,W,hd] is ever materialised; Q/K/V are
// read in place from the packed qkv GEMM output and O is written in the [B,N,H*hd] layout the output
// projection consumes.
//
// The op is HBM-bound at W=7 (AI ~ W/2 FLOP/B), so the kernels are SIMT with 16/32-byte
// vector loads, fp32 math, exp2 with pre-scaled logits, and warp-shuffle reductions inside the
// hd/8-lane group that owns a query (or a key, in the dK/dV pass).  Gradients are gathered per key —
// no atomics, deterministic.
#include <math_constants.h>

#include "favit_common.cuh"

namespace favit {

// whole-sequence TMA path (mhla_window_attn_seq.cu): head_dim 64, N <= 400 (the forward only where the chunked one does not apply)
bool attn_seq_applicable(int hd, int window, int N, favit_dtype dtype, const uint8_t* mask, const void* q, const void* k,
                         const void* v, int64_t sb, int64_t sn, int64_t shh);
int attn_seq_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int window,
                 float scale, int64_t sb, int64_t sn, int64_t shh, cudaStream_t st);
int attn_seq_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout, void* dq,
                 void* dk, void* dv, float* colsum, int B, int H, int N, int window, float scale, int64_t sb, int64_t sn,
                 int64_t shh, cudaStream_t st);
// chunked TMA path (mhla_window_attn_chunk.cu): head_dim 64, window <= 15, N > 48 (forward) / N > 400 (backward)
bool attn_chunk_applicable(int hd, int window, int N, int B, int H, favit_dtype dtype, const uint8_t* mask, const void* q,
                           const void* k, const void* v, int64_t sb, int64_t sn, int64_t shh, bool backward);
int attn_chunk_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int window,
                   float scale, int64_t sb, int64_t sn, int64_t shh, cudaStream_t st);
int attn_chunk_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout, void* dq,
                   void* dk, void* dv, float* colsum, int B, int H, int N, int window, float scale, int64_t sb, int64_t sn,
                   int64_t shh, cudaStream_t st);
// tensor-core tile path (mhla_window_attn_mma.cu)
bool attn_mma_applicable(int hd, int window, favit_dtype dtype, const uint8_t* mask);
// mhla_window_attn_tc.cu: tcgen05 / TMEM forward for wide windows (17 <= W <= 65, head_dim 64, bf16, N >= W)
bool attn_tc_applicable(int hd, int window, int N, favit_dtype dtype, const uint8_t* mask, const void* q, const void* k,
                        const void* v, int64_t sb, int64_t sn, int64_t shh);
int attn_tc_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int window,
                float scale, int64_t sb, int64_t sn, int64_t shh, cudaStream_t st);
int attn_tc_bwd(const void* q, const void* k, const void* v, const void* out, const float* lse, const void* dout, void* dq,
                void* dk, void* dv, float* delta, int B, int H, int N, int window, float scale, int64_t sb, int64_t sn,
                int64_t shh, cudaStream_t st);
int attn_mma_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int hd,
                 int window, float scale, int64_t sb, int64_t sn, int64_t shh, cudaStream_t st);
int attn_mma_bwd(const void* q, const void* k, const void* v, const void* o, const float* lse, const void* dout,
                 void* dq, void* dk, void* dv, float* delta, float* colsum, int B, int H, int N, int hd, int window,
                 float scale, int64_t sb, int64_t sn, int64_t shh, cudaStream_t st);

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct AttnShape {
  int B, H, N, W;
  int64_t sb, sn, sh;  // q/k/v strides in elements
  float scale_log2;    // scale * log2(e)
  float scale;
  float drop_p = 0.f;      // attention-probability dropout (mhla.py:147); 0 = off
  float inv_keep = 1.f;    // 1 / (1 - drop_p)
  uint64_t seed = 0;
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

// multiplicity of key j in query i's window
__device__ __forceinline__ int window_mult(const WindowRow& r, int j) {
  return ((j >= r.s && j < r.e) ? 1 : 0) + ((j == r.tgt) ? r.pad : 0);
}

// ---- attention-probability dropout (mhla.py:147, nn.Dropout on the softmax output [B,H,N,W]) ---------------------
// The reference draws one Bernoulli per WINDOW SLOT, so the pad copies of an edge key are dropped independently.  Slot
// `pos` of row (b, h, i) is kept iff a counter-based hash of (seed, ((b*H + h)*N + i)*W + pos) gives u >= p (splitmix64
// finaliser, 24-bit uniform): stateless, so forward, the dQ pass and the dK/dV pass regenerate the same mask, and the
// CPU oracle reproduces it bit for bit.  Slot order is the reference's (mhla.py:72-79): s == 0 rows hold the band
// first and the copies of key N-1 after it, the other rows the copies of key 0 first.
__device__ __forceinline__ bool drop_keep(uint64_t seed, uint64_t counter, float p) {
  uint64_t z = seed + counter * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.f / 16777216.f) >= p;
}
// sum over the copies of key j in row i of keep / (1 - p): replaces the multiplicity in everything downstream of the
// softmax normaliser
__device__ __forceinline__ float kept_weight(const AttnShape& sh, int b, int h, int i, const WindowRow& r, int j) {
  const uint64_t row = (((uint64_t)b * sh.H + h) * sh.N + i) * (uint64_t)sh.W;
  const int band = r.e - r.s;
  int kept = 0;
  if (j >= r.s && j < r.e) kept += drop_keep(sh.seed, row + (r.s == 0 ? j - r.s : r.pad + j - r.s), sh.drop_p) ? 1 : 0;
  if (j == r.tgt) {
    const int p0 = (r.s == 0) ? band : 0;
    for (int t = 0; t < r.pad; ++t) kept += drop_keep(sh.seed, row + p0 + t, sh.drop_p) ? 1 : 0;
  }
  return (float)kept * sh.inv_keep;
}

// Sum over the LPQ consecutive lanes that own one query/key.  The shuffle names only the lanes of the
// group, so groups of one warp may diverge from each other (different window lengths at the edges).
template <int LPQ>
__device__ __forceinline__ float group_sum(float v, unsigned gmask) {
#pragma unroll
  for (int o = LPQ / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

template <int LPQ>
__device__ __forceinline__ unsigned group_mask() {
  const unsigned lane = threadIdx.x & 31u;
  return (LPQ >= 32) ? 0xffffffffu : (((1u << LPQ) - 1u) << (lane & ~(unsigned)(LPQ - 1)));
}

template <typename T, int HD>
__global__ void __launch_bounds__(256) attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                       const T* __restrict__ v,
                                                       const uint8_t* __restrict__ mask,
                                                       T* __restrict__ out, float* __restrict__ lse,
                                                       AttnShape sh) {
  constexpr int LPQ = HD / 8;
  const int64_t total = (int64_t)sh.B * sh.H * sh.N;
  int64_t g = (int64_t)blockIdx.x * (blockDim.x / LPQ) + threadIdx.x / LPQ;
  const bool valid = g < total;
  if (!valid) g = total - 1;
  const int sub = threadIdx.x % LPQ;
  const unsigned gmask = group_mask<LPQ>();
  const int i = (int)(g % sh.N);
  const int h = (int)((g / sh.N) % sh.H);
  const int b = (int)(g / ((int64_t)sh.N * sh.H));
  const int64_t base = (int64_t)b * sh.sb + (int64_t)h * sh.sh + sub * 8;

  float qf[8];
  load8(q + base + (int64_t)i * sh.sn, qf);
#pragma unroll
  for (int d = 0; d < 8; ++d) qf[d] *= sh.scale_log2;

  const WindowRow r = window_row(i, sh.N, sh.W);
  const int band = r.e - r.s;
  const bool extra = r.pad > 0 && !(r.tgt >= r.s && r.tgt < r.e);
  const int cnt = band + (extra ? 1 : 0);
  const uint8_t* mrow = mask ? mask + ((int64_t)b * sh.N + i) * sh.N : nullptr;

  float m_run = -CUDART_INF_F, l_run = 0.f;
  float acc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = 0.f;

  constexpr int JB = 4;
  for (int t0 = 0; t0 < cnt; t0 += JB) {
    float sc[JB];
    float wf[JB];  // dropout: kept copies / ((1 - p) * multiplicity); 1 without dropout
    float vf[JB][8];
#pragma unroll
    for (int u = 0; u < JB; ++u) {
      const int t = t0 + u;
      sc[u] = -CUDART_INF_F;
      wf[u] = 1.f;
      if (t < cnt) {
        const int j = (t < band) ? (r.s + t) : r.tgt;
        const int mult = (t < band) ? (1 + ((j == r.tgt) ? r.pad : 0)) : r.pad;
        if (sh.drop_p > 0.f) wf[u] = kept_weight(sh, b, h, i, r, j) / (float)mult;
        float kf[8];
        load8(k + base + (int64_t)j * sh.sn, kf);
        load8(v + base + (int64_t)j * sh.sn, vf[u]);
        float dot = 0.f;
#pragma unroll
        for (int d = 0; d < 8; ++d) dot = fmaf(qf[d], kf[d], dot);
        dot = group_sum<LPQ>(dot, gmask);
        const bool keep = (mrow == nullptr) || (mrow[j] != 0);
        if (keep) sc[u] = dot + ((mult > 1) ? log2f((float)mult) : 0.f);
      } else {
#pragma unroll
        for (int d = 0; d < 8; ++d) vf[u][d] = 0.f;
      }
    }
    float m_new = m_run;
#pragma unroll
    for (int u = 0; u < JB; ++u) m_new = fmaxf(m_new, sc[u]);
    if (m_new != -CUDART_INF_F) {
      const float corr = exp2f(m_run - m_new);  // m_run = -inf -> 0
      l_run *= corr;
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] *= corr;
#pragma unroll
      for (int u = 0; u < JB; ++u) {
        const float p = exp2f(sc[u] - m_new);  // -inf -> 0
        l_run += p;                            // the softmax normaliser never sees the dropout
        const float pw = p * wf[u];
#pragma unroll
        for (int d = 0; d < 8; ++d) acc[d] = fmaf(pw, vf[u][d], acc[d]);
      }
      m_run = m_new;
    }
  }
  if (valid) {
    const float inv = 1.f / l_run;  // fully masked row: 0 * inf -> NaN, like softmax of all -inf
    float of[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) of[d] = acc[d] * inv;
    store8(out + (((int64_t)b * sh.N + i) * sh.H + h) * HD + sub * 8, of);
    if (sub == 0) lse[((int64_t)b * sh.H + h) * sh.N + i] = (m_run + log2f(l_run)) * kLn2;
  }
}

// dQ pass (one group per query); also emits delta_i = dO_i . O_i for the dK/dV pass.
template <typename T, int HD>
__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(
    const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
    const uint8_t* __restrict__ mask, const T* __restrict__ o, const float* __restrict__ lse,
    const T* __restrict__ dout, T* __restrict__ dq, float* __restrict__ delta, AttnShape sh) {
  constexpr int LPQ = HD / 8;
  const int64_t total = (int64_t)sh.B * sh.H * sh.N;
  int64_t g = (int64_t)blockIdx.x * (blockDim.x / LPQ) + threadIdx.x / LPQ;
  const bool valid = g < total;
  if (!valid) g = total - 1;
  const int sub = threadIdx.x % LPQ;
  const unsigned gmask = group_mask<LPQ>();
  const int i = (int)(g % sh.N);
  const int h = (int)((g / sh.N) % sh.H);
  const int b = (int)(g / ((int64_t)sh.N * sh.H));
  const int64_t base = (int64_t)b * sh.sb + (int64_t)h * sh.sh + sub * 8;
  const int64_t obase = (((int64_t)b * sh.N + i) * sh.H + h) * HD + sub * 8;

  float qf[8], dof[8], of[8];
  load8(q + base + (int64_t)i * sh.sn, qf);
  load8(dout + obase, dof);
  load8(o + obase, of);
  float dl = 0.f;
#pragma unroll
  for (int d = 0; d < 8; ++d) dl = fmaf(dof[d], of[d], dl);
  dl = group_sum<LPQ>(dl, gmask);
  const float L2 = lse[((int64_t)b * sh.H + h) * sh.N + i] * kLog2e;

  const WindowRow r = window_row(i, sh.N, sh.W);
  const int band = r.e - r.s;
  const bool extra = r.pad > 0 && !(r.tgt >= r.s && r.tgt < r.e);
  const int cnt = band + (extra ? 1 : 0);
  const uint8_t* mrow = mask ? mask + ((int64_t)b * sh.N + i) * sh.N : nullptr;

  float acc[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) acc[d] = 0.f;

  for (int t = 0; t < cnt; ++t) {
    const int j = (t < band) ? (r.s + t) : r.tgt;
    const int mult = (t < band) ? (1 + ((j == r.tgt) ? r.pad : 0)) : r.pad;
    float kf[8], vf[8];
    load8(k + base + (int64_t)j * sh.sn, kf);
    load8(v + base + (int64_t)j * sh.sn, vf);
    float dot = 0.f, dp = 0.f;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      dot = fmaf(qf[d], kf[d], dot);
      dp = fmaf(dof[d], vf[d], dp);
    }
    dot = group_sum<LPQ>(dot, gmask);
    dp = group_sum<LPQ>(dp, gmask);
    const bool keep = (mrow == nullptr) || (mrow[j] != 0);
    if (keep) {
      const float p = exp2f(dot * sh.scale_log2 + ((mult > 1) ? log2f((float)mult) : 0.f) - L2);
      const float wfac = sh.drop_p > 0.f ? kept_weight(sh, b, h, i, r, j) / (float)mult : 1.f;
      const float ds = p * (wfac * dp - dl);
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] = fmaf(ds, kf[d], acc[d]);
    }
  }
  if (valid) {
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] *= sh.scale;
    store8(dq + base + (int64_t)i * sh.sn, acc);
    if (sub == 0) delta[((int64_t)b * sh.H + h) * sh.N + i] = dl;
  }
}

// dK/dV pass (one group per key): gathers from every query whose window holds this key.
template <typename T, int HD>
__global__ void __launch_bounds__(256) attn_bwd_dkv_kernel(
    const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
    const uint8_t* __restrict__ mask, const float* __restrict__ lse, const T* __restrict__ dout,
    const float* __restrict__ delta, T* __restrict__ dk, T* __restrict__ dv, AttnShape sh) {
  constexpr int LPQ = HD / 8;
  const int64_t total = (int64_t)sh.B * sh.H * sh.N;
  int64_t g = (int64_t)blockIdx.x * (blockDim.x / LPQ) + threadIdx.x / LPQ;
  const bool valid = g < total;
  if (!valid) g = total - 1;
  const int sub = threadIdx.x % LPQ;
  const unsigned gmask = group_mask<LPQ>();
  const int j = (int)(g % sh.N);
  const int h = (int)((g / sh.N) % sh.H);
  const int b = (int)(g / ((int64_t)sh.N * sh.H));
  const int64_t base = (int64_t)b * sh.sb + (int64_t)h * sh.sh + sub * 8;
  const int N = sh.N, W = sh.W, hw = W >> 1;

  float kf[8], vf[8];
  load8(k + base + (int64_t)j * sh.sn, kf);
  load8(v + base + (int64_t)j * sh.sn, vf);
  float dkf[8], dvf[8];
#pragma unroll
  for (int d = 0; d < 8; ++d) { dkf[d] = 0.f; dvf[d] = 0.f; }

  // three candidate ranges: the band, then the duplicated-edge contributors of key N-1 / key 0
  const int lo0 = max(0, j - hw), hi0 = min(N - 1, j + hw);
  int lo[3] = {lo0, 0, 0}, hi[3] = {hi0, -1, -1};
  if (j == N - 1) { lo[1] = 0; hi[1] = min(min(hw, N - 1), lo0 - 1); }
  if (j == 0) { lo[2] = max(max(hw + 1, N - hw), hi0 + 1); hi[2] = N - 1; }

  const float* lse_bh = lse + ((int64_t)b * sh.H + h) * N;
  const float* dl_bh = delta + ((int64_t)b * sh.H + h) * N;
#pragma unroll
  for (int rg = 0; rg < 3; ++rg) {
    for (int i = lo[rg]; i <= hi[rg]; ++i) {
      const WindowRow r = window_row(i, N, W);
      const int mult = window_mult(r, j);
      if (mult == 0) continue;
      if (mask && mask[((int64_t)b * N + i) * N + j] == 0) continue;
      float qf[8], dof[8];
      load8(q + base + (int64_t)i * sh.sn, qf);
      load8(dout + (((int64_t)b * N + i) * sh.H + h) * HD + sub * 8, dof);
      float dot = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        dot = fmaf(qf[d], kf[d], dot);
        dp = fmaf(dof[d], vf[d], dp);
      }
      dot = group_sum<LPQ>(dot, gmask);
      dp = group_sum<LPQ>(dp, gmask);
      const float p = exp2f(dot * sh.scale_log2 + ((mult > 1) ? log2f((float)mult) : 0.f) -
                            lse_bh[i] * kLog2e);
      const float wfac = sh.drop_p > 0.f ? kept_weight(sh, b, h, i, r, j) / (float)mult : 1.f;
      const float ds = p * (wfac * dp - dl_bh[i]);
      const float pv = p * wfac;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        dvf[d] = fmaf(pv, dof[d], dvf[d]);
        dkf[d] = fmaf(ds, qf[d], dkf[d]);
      }
    }
  }
  if (valid) {
#pragma unroll
    for (int d = 0; d < 8; ++d) dkf[d] *= sh.scale;
    store8(dk + base + (int64_t)j * sh.sn, dkf);
    store8(dv + base + (int64_t)j * sh.sn, dvf);
  }
}

template <typename T, int HD>
int launch_fwd(const void* q, const void* k, const void* v, const uint8_t* mask, void* out, float* lse,
               const AttnShape& sh, cudaStream_t st) {
  constexpr int LPQ = HD / 8;
  const int64_t total = (int64_t)sh.B * sh.H * sh.N;
  const int threads = 256;
  const int64_t blocks = ceil_div64(total, threads / LPQ);
  attn_fwd_kernel<T, HD><<<(unsigned)blocks, threads, 0, st>>>((const T*)q, (const T*)k, (const T*)v, mask,
                                                               (T*)out, lse, sh);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

template <typename T, int HD>
int launch_bwd(const void* q, const void* k, const void* v, const uint8_t* mask, const void* o,
               const float* lse, const void* dout, void* dq, void* dk, void* dv, float* delta,
               const AttnShape& sh, cudaStream_t st) {
  constexpr int LPQ = HD / 8;
  const int64_t total = (int64_t)sh.B * sh.H * sh.N;
  const int threads = 256;
  const int64_t blocks = ceil_div64(total, threads / LPQ);
  attn_bwd_dq_kernel<T, HD><<<(unsigned)blocks, threads, 0, st>>>(
      (const T*)q, (const T*)k, (const T*)v, mask, (const T*)o, lse, (const T*)dout, (T*)dq, delta, sh);
  FAVIT_CHECK_LAUNCH();
  attn_bwd_dkv_kernel<T, HD><<<(unsigned)blocks, threads, 0, st>>>(
      (const T*)q, (const T*)k, (const T*)v, mask, lse, (const T*)dout, delta, (T*)dk, (T*)dv, sh);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

int check_common(const void* q, const void* k, const void* v, int B, int H, int N, int hd, int window,
                 int64_t sb, int64_t sn, int64_t shh, favit_dtype dtype, float dropout_p) {
  FAVIT_CHECK_ARG(q && k && v, "mhla_attn: null q/k/v");
  FAVIT_CHECK_ARG(B > 0 && H > 0 && N > 0, "mhla_attn: B,H,N must be positive (got %d,%d,%d)", B, H, N);
  FAVIT_CHECK_ARG(window >= 1, "mhla_attn: window must be >= 1");
  FAVIT_CHECK_ARG(!(window % 2 == 0 && N > window),
                  "mhla_attn: even window_size=%d with seq_len=%d > window is ragged in the reference "
                  "(models/mhla.py:83)", window, N);
  FAVIT_CHECK_ARG(dtype == FAVIT_F32 || dtype == FAVIT_BF16, "mhla_attn: bad dtype");
  if (!(hd == 16 || hd == 32 || hd == 64 || hd == 128)) {
    set_error("mhla_attn: head_dim %d unsupported (16/32/64/128)", hd);
    return FAVIT_ERR_UNSUPPORTED;
  }
  const int64_t al = (dtype == FAVIT_BF16) ? 8 : 4;  // elements per 16 bytes
  FAVIT_CHECK_ARG(sb % al == 0 && sn % al == 0 && shh % al == 0, "mhla_attn: strides must keep 16-byte alignment");
  FAVIT_CHECK_ARG(((uintptr_t)q % 16 == 0) && ((uintptr_t)k % 16 == 0) && ((uintptr_t)v % 16 == 0),
                  "mhla_attn: q/k/v must be 16-byte aligned");
  FAVIT_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "mhla_attn: dropout_p must be in [0, 1) (got %g)", (double)dropout_p);
  return FAVIT_OK;
}

}  // namespace
}  // namespace favit

using namespace favit;

#define FAVIT_DISPATCH_HD(T, fn, ...)                  \
  switch (hd) {                                        \
    case 16: return fn<T, 16>(__VA_ARGS__);            \
    case 32: return fn<T, 32>(__VA_ARGS__);            \
    case 64: return fn<T, 64>(__VA_ARGS__);            \
    default: return fn<T, 128>(__VA_ARGS__);           \
  }

extern "C" int favit_mhla_attn_fwd(const void* q, const void* k, const void* v, const uint8_t* mask, void* out,
                                   float* lse, int B, int H, int N, int hd, int window, float scale,
                                   int64_t stride_b, int64_t stride_n, int64_t stride_h, favit_dtype dtype,
                                   float dropout_p, uint64_t seed, favit_stream stream) {
  int rc = check_common(q, k, v, B, H, N, hd, window, stride_b, stride_n, stride_h, dtype, dropout_p);
  if (rc) return rc;
  FAVIT_CHECK_ARG(out && lse, "mhla_attn_fwd: null out/lse");
  const bool drop = dropout_p > 0.f;  // dropout (training, non-default) runs on the general kernels
  if (!drop && attn_chunk_applicable(hd, window, N, B, H, dtype, mask, q, k, v, stride_b, stride_n, stride_h, false) &&
      ((uintptr_t)out % 16) == 0) {
    note_kernel("attn_chunk_fwd (TMA sequence chunks, mma.sync)");
    return attn_chunk_fwd(q, k, v, out, lse, B, H, N, window, scale, stride_b, stride_n, stride_h, (cudaStream_t)stream);
  }
  if (!drop && attn_seq_applicable(hd, window, N, dtype, mask, q, k, v, stride_b, stride_n, stride_h) &&
      ((uintptr_t)out % 16) == 0) {
    note_kernel("attn_seq_fwd (TMA whole-sequence, mma.sync)");
    return attn_seq_fwd(q, k, v, out, lse, B, H, N, window, scale, stride_b, stride_n, stride_h, (cudaStream_t)stream);
  }
  if (!drop && attn_tc_applicable(hd, window, N, dtype, mask, q, k, v, stride_b, stride_n, stride_h) &&
      ((uintptr_t)out % 16) == 0) {
    note_kernel("attn_tc_fwd (tcgen05 / TMEM, 128-query tiles)");
    return attn_tc_fwd(q, k, v, out, lse, B, H, N, window, scale, stride_b, stride_n, stride_h, (cudaStream_t)stream);
  }
  if (!drop && attn_mma_applicable(hd, window, dtype, mask)) {
    note_kernel("attn_mma_fwd (per-warp staging, mma.sync)");
    return attn_mma_fwd(q, k, v, out, lse, B, H, N, hd, window, scale, stride_b, stride_n, stride_h,
                        (cudaStream_t)stream);
  }
  note_kernel("attn_simt_fwd");
  AttnShape sh{B, H, N, window, stride_b, stride_n, stride_h, scale * kLog2e, scale};
  if (drop) { sh.drop_p = dropout_p; sh.inv_keep = 1.f / (1.f - dropout_p); sh.seed = seed; }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == FAVIT_BF16) {
    FAVIT_DISPATCH_HD(__nv_bfloat16, launch_fwd, q, k, v, mask, out, lse, sh, st)
  } else {
    FAVIT_DISPATCH_HD(float, launch_fwd, q, k, v, mask, out, lse, sh, st)
  }
}

extern "C" int favit_colsum(const void* x, favit_dtype dtype, float* out, int M, int N, int64_t ld, favit_stream stream);

extern "C" int favit_mhla_attn_bwd(const void* q, const void* k, const void* v, const uint8_t* mask,
                                   const void* out, const float* lse, const void* dout, void* dq, void* dk,
                                   void* dv, float* delta, float* dqkv_colsum, int B, int H, int N, int hd, int window,
                                   float scale,
                                   int64_t stride_b, int64_t stride_n, int64_t stride_h, favit_dtype dtype,
                                   float dropout_p, uint64_t seed, favit_stream stream) {
  int rc = check_common(q, k, v, B, H, N, hd, window, stride_b, stride_n, stride_h, dtype, dropout_p);
  if (rc) return rc;
  FAVIT_CHECK_ARG(out && lse && dout && dq && dk && dv && delta, "mhla_attn_bwd: null pointer");
  const bool drop = dropout_p > 0.f;
  if (!drop && attn_chunk_applicable(hd, window, N, B, H, dtype, mask, q, k, v, stride_b, stride_n, stride_h, true) &&
      ((uintptr_t)dout % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dq % 16) == 0 && ((uintptr_t)dk % 16) == 0 &&
      ((uintptr_t)dv % 16) == 0) {
    note_kernel("attn_chunk_bwd (TMA sequence chunks with halo tiles, mma.sync)");
    return attn_chunk_bwd(q, k, v, out, lse, dout, dq, dk, dv, dqkv_colsum, B, H, N, window, scale, stride_b, stride_n,
                          stride_h, (cudaStream_t)stream);
  }
  if (!drop && attn_seq_applicable(hd, window, N, dtype, mask, q, k, v, stride_b, stride_n, stride_h) &&
      ((uintptr_t)dout % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dq % 16) == 0 && ((uintptr_t)dk % 16) == 0 &&
      ((uintptr_t)dv % 16) == 0) {
    note_kernel("attn_seq_bwd (TMA whole-sequence, mma.sync)");
    return attn_seq_bwd(q, k, v, out, lse, dout, dq, dk, dv, dqkv_colsum, B, H, N, window, scale, stride_b, stride_n,
                        stride_h, (cudaStream_t)stream);
  }
  if (!drop && attn_tc_applicable(hd, window, N, dtype, mask, q, k, v, stride_b, stride_n, stride_h) &&
      ((uintptr_t)dout % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)dq % 16) == 0 && ((uintptr_t)dk % 16) == 0 &&
      ((uintptr_t)dv % 16) == 0) {
    note_kernel("attn_tc_bwd (tcgen05 / TMEM: dQ query-major, dK dV key-major, edge rows)");
    rc = attn_tc_bwd(q, k, v, out, lse, dout, dq, dk, dv, delta, B, H, N, window, scale, stride_b, stride_n, stride_h,
                     (cudaStream_t)stream);
    if (rc || !dqkv_colsum) return rc;
    FAVIT_CHECK_ARG(stride_b == (int64_t)N * stride_n, "mhla_attn_bwd: dqkv_colsum needs batch-contiguous dq/dk/dv");
    const void* parts_tc[3] = {dq, dk, dv};
    for (int t = 0; t < 3; ++t) {
      rc = favit_colsum(parts_tc[t], dtype, dqkv_colsum + (size_t)t * H * hd, B * N, H * hd, stride_n, stream);
      if (rc) return rc;
    }
    return FAVIT_OK;
  }
  if (!drop && attn_mma_applicable(hd, window, dtype, mask)) {
    note_kernel("attn_mma_bwd (per-warp staging, mma.sync)");
    return attn_mma_bwd(q, k, v, out, lse, dout, dq, dk, dv, delta, dqkv_colsum, B, H, N, hd, window, scale, stride_b,
                        stride_n, stride_h, (cudaStream_t)stream);
  }
  note_kernel("attn_simt_bwd");
  AttnShape sh{B, H, N, window, stride_b, stride_n, stride_h, scale * kLog2e, scale};
  if (drop) { sh.drop_p = dropout_p; sh.inv_keep = 1.f / (1.f - dropout_p); sh.seed = seed; }
  cudaStream_t st = (cudaStream_t)stream;
  auto run = [&]() -> int {
    if (dtype == FAVIT_BF16) {
      FAVIT_DISPATCH_HD(__nv_bfloat16, launch_bwd, q, k, v, mask, out, lse, dout, dq, dk, dv, delta, sh, st)
    } else {
      FAVIT_DISPATCH_HD(float, launch_bwd, q, k, v, mask, out, lse, dout, dq, dk, dv, delta, sh, st)
    }
  };
  rc = run();
  if (rc || !dqkv_colsum) return rc;
  // general path: the column sums (bias gradient of the qkv projection) take one more pass over dq, dk, dv
  FAVIT_CHECK_ARG(stride_b == (int64_t)N * stride_n, "mhla_attn_bwd: dqkv_colsum needs batch-contiguous dq/dk/dv");
  const void* parts[3] = {dq, dk, dv};
  for (int t = 0; t < 3; ++t) {
    rc = favit_colsum(parts[t], dtype, dqkv_colsum + (size_t)t * H * hd, B * N, H * hd, stride_n, stream);
    if (rc) return rc;
  }
  return FAVIT_OK;
}
