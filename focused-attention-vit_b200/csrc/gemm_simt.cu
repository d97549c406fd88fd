// fp32 SIMT GEMM for the parity path of the MHLA linears (sm_100a).
//
// The reference runs nn.Linear in true fp32 when autocast is off (torch keeps TF32 disabled for matmul by
// default), and BASELINE.json asks for ~1e-4 relative agreement there.  kind::tf32 tensor-core MMAs carry a
// 10-bit mantissa and would not meet that bar, so the fp32 path uses plain FFMA tiles: 64x64x16 CTA tile,
// 4x4 register micro-tile, shared-memory staging with a layout chosen so that global reads are coalesced for
// either operand orientation.  It is the correctness path; the throughput path is gemm_tcgen05.cu (bf16).
//
//   C[m,n] = sum_k A(m,k) * B(n,k),   A(m,k) = A[m*sam + k*sak],   B(n,k) = B[n*sbn + k*sbk]
#include "favit_common.cuh"
#include "gemm_epilogue.cuh"
#include "gemm_tcgen05.h"

namespace favit {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtArgs {
  const float* A; const float* B; float* C;
  const float* bias; const float* aux; float* aux_out;
  const float* residual; int64_t ldres;
  int M, N, K;
  int64_t sam, sak, sbn, sbk, ldc;
  int epilogue;
  int atomic;   // split-K: accumulate with atomicAdd
  int k_per_split;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtArgs a) {
  __shared__ float sA[TK][TM + 4];
  __shared__ float sB[TK][TN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kbeg = blockIdx.z * a.k_per_split;
  const int kend = min(a.K, kbeg + a.k_per_split);
  const int tx = tid % 16, ty = tid / 16;  // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kcontig = (a.sak == 1);
  const bool b_kcontig = (a.sbk == 1);
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    // stage A tile [TM x TK] and B tile [TN x TK]; 1024 elements each, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int r, kk;
      if (a_kcontig) { r = idx / TK; kk = idx % TK; } else { kk = idx / TM; r = idx % TM; }
      const int gm = m0 + r, gk = k0 + kk;
      sA[kk][r] = (gm < a.M && gk < kend) ? a.A[(int64_t)gm * a.sam + (int64_t)gk * a.sak] : 0.f;
      if (b_kcontig) { r = idx / TK; kk = idx % TK; } else { kk = idx / TN; r = idx % TN; }
      const int gn = n0 + r;
      const int gk2 = k0 + kk;
      sB[kk][r] = (gn < a.N && gk2 < kend) ? a.B[(int64_t)gn * a.sbn + (int64_t)gk2 * a.sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      const float af[4] = {av.x, av.y, av.z, av.w};
      const float bf[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      const int64_t off = (int64_t)gm * a.ldc + gn;
      float v = acc[i][j];
      if (a.atomic) {
        atomicAdd(a.C + off, v);
        continue;
      }
      if (a.bias) v += a.bias[gn];
      if (a.epilogue == FAVIT_EPI_GELU) {
        if (a.aux_out) a.aux_out[off] = v;
        v = gelu_erf(v);
      } else if (a.epilogue == FAVIT_EPI_DGELU_MUL) {
        v *= dgelu_erf(a.aux[off]);
      }
      if (a.residual) v += a.residual[(int64_t)gm * a.ldres + gn];
      a.C[off] = v;
    }
  }
}

}  // namespace

int gemm_simt_launch(const float* A, const float* B, float* C, const float* bias, const float* aux,
                     float* aux_out, const float* residual, int64_t ldres, int M, int N, int K, int64_t sam,
                     int64_t sak, int64_t sbn, int64_t sbk, int64_t ldc, int epilogue, int splits, int accumulate,
                     cudaStream_t st) {
  SimtArgs a{A, B, C, bias, aux, aux_out, residual, ldres, M, N, K, sam, sak, sbn, sbk, ldc, epilogue,
             (splits > 1 || accumulate) ? 1 : 0, K};
  if (splits > 1) {
    int kps = ceil_div(ceil_div(K, splits), TK) * TK;
    a.k_per_split = kps;
    splits = ceil_div(K, kps);
  }
  dim3 grid(ceil_div(N, TN), ceil_div(M, TM), splits > 1 ? splits : 1);
  note_kernel("gemm_simt_kernel splits=%d", splits);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(a);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

}  // namespace favit
