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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
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
    const bool ok = (unsigned)(k - lo) <= (unsigned)(hi - lo);
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
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

}  // namespace favit
