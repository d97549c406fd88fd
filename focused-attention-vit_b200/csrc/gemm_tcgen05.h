// Internal interface of the tcgen05 bf16 GEMM (gemm_tcgen05.cu) and the fp32 SIMT GEMM (gemm_simt.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_epilogue.cuh"

namespace favit {
namespace tc {

// What happens to an fp32 accumulator tile on its way out of TMEM.
//   v = acc (+ bias[col]);  act: GELU (optionally saving the pre-activation), or v *= gelu'(aux);
//   v *= dropout keep-mask / (1 - p);  v += residual[row,col];  C[row,col] = v   (or atomically C += v when accumulate / split-K).
struct Epilogue {
  void* c = nullptr;
  int64_t ldc = 0;
  int c_dtype = 0;                // favit_dtype of C
  const float* bias = nullptr;    // [N] fp32
  const void* residual = nullptr; // [M,N]
  int64_t ldres = 0;
  int res_dtype = 0;
  void* aux_out = nullptr;        // [M,N] bf16 pre-activation written when act == GELU
  const void* aux = nullptr;      // [M,N] bf16 pre-activation read when act == DGELU_MUL
  int64_t ldaux = 0;
  int act = 0;                    // favit_epilogue
  int accumulate = 0;             // C += result (fp32 C only)
  int split_ok = 0;               // caller allows split-K (C is zeroed or being accumulated into)
  float* colsum = nullptr;        // [N] fp32, accumulated: column sums of a bf16 C (CTA-pair kernel only)
  DropSpec drop;                  // dropout applied after `act`, before the residual (drop.seed == nullptr: none)
};

// C[M,N] = A.B^T with the operand storage flags described in gemm_tcgen05.cu.
int gemm_bf16(const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, int M, int N, int K,
              const Epilogue& epi, int force_bn, int force_splits, cudaStream_t st);

// CTA-pair kernel (gemm_tcgen05_2cta.cu): 256 x 256 tiles, TMA-store epilogue.
bool gemm_bf16_2cta_applicable(int M, int N, int K, const Epilogue& e, int64_t lda, int64_t ldb);
int gemm_bf16_2cta(const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, int M, int N, int K,
                   const Epilogue& epi, int force_splits, cudaStream_t st);

// How the CTA-pair kernel hands out its tiles: 0 = static striding, 1 = work stealing; any other value only queries.
// Returns the mode in force.
int gemm_tile_scheduler(int mode);

}  // namespace tc

// fp32 SIMT GEMM (parity path): C[m,n] = sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk]
int gemm_simt_launch(const float* A, const float* B, float* C, const float* bias, const float* aux, float* aux_out,
                     const float* residual, int64_t ldres, int M, int N, int K, int64_t sam, int64_t sak, int64_t sbn,
                     int64_t sbk, int64_t ldc, int epilogue, int splits, int accumulate, cudaStream_t st);

}  // namespace favit
