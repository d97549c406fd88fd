// Element-wise epilogue math shared by the SIMT and tcgen05 GEMMs.
#pragma once
#include <cuda_runtime.h>

namespace favit {

// nn.GELU() default (erf form), reference models/vit.py:120, models/mhla.py:199 — exact erff, used by the fp32 path.
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// bf16 path: Phi(x) = 0.5 (1 + erf(x / sqrt 2)) as a logistic of an odd polynomial,
//   Phi(x) ~= 1 / (1 + exp(-(c1 x + c3 x^3 + c5 x^5))),   fitted on [-9, 9]:
//   |Phi error| <= 5.7e-5, |gelu error| <= 2.9e-5 absolute — below bf16 resolution for every |y| >= 0.008, and the
//   outputs of these epilogues are rounded to bf16 anyway.  Round 1 evaluated the logistic with EX2 + RCP (2 MUFU + 6
//   FMA-class instructions per element instead of erff's ~25); round 2 uses the tanh identity below (1 MUFU): the GELU epilogue has to drain a 128 x 256 tile faster than the tensor pipe fills the next one
//   (12 k-blocks = 6144 cycles at K = 768), and at 17 instructions per element it did not.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigma(k) = 0.5 + 0.5 tanh(k / 2) exactly, so the logistic fit costs ONE special-function instruction (MUFU.TANH)
// instead of two (EX2 + RCP): the GELU / GELU' epilogues are SFU-bound (a 128 x 256 tile = 64 Ki MUFU at 16 per clock per
// SM against 6 144 tensor-pipe cycles at K = 768).  Coefficients = those of k(x) / 2:  c1 x + c3 x^3 + c5 x^5 with
// c1 = 0.79746547 (sqrt(2 / pi) = 0.79788), c3 = 3.7043716e-2, c5 = -3.5764635e-4.  tanh.approx has a relative error of
// 2^-11: |Phi error| <= 2.5e-4, an order of magnitude below bf16 resolution of the outputs these epilogues store.
__device__ __forceinline__ float phi_half_arg(float x) {
  // the fit is monotone on [-8, 8]; clamping x^2 (one instruction) keeps the argument monotone beyond: Phi -> 0 / 1
  const float x2 = fminf(x * x, 64.f);
  float p = fmaf(-3.5764635e-4f, x2, 3.7043716e-2f);
  p = fmaf(p, x2, 0.79746547f);
  return p * x;
}
__device__ __forceinline__ float phi_fast(float x) { return fmaf(0.5f, tanh_approx(phi_half_arg(x)), 0.5f); }
__device__ __forceinline__ float gelu_fast(float x) { return x * phi_fast(x); }
// gelu'(x) = Phi(x) + x * phi(x).  With Phi = sigma(k(x)) the density is Phi' = Phi (1 - Phi) k'(x) = (1 - t^2) k'(x) / 4,
// t = tanh(k / 2): the exact derivative of gelu_fast, still one special-function instruction;
// |error| vs the erf form <= 3e-4.  k'(x) = 2 (c1 + 3 c3 x^2 + 5 c5 x^4).
__device__ __forceinline__ float dgelu_fast(float x) {
  const float x2 = fminf(x * x, 64.f);
  float p = fmaf(-3.5764635e-4f, x2, 3.7043716e-2f);
  p = fmaf(p, x2, 0.79746547f);
  const float t = tanh_approx(p * x);
  float kq = fmaf(-8.9411588e-4f, x2, 5.5565574e-2f);   // k'(x) / 4 = (c1 + 3 c3 x^2 + 5 c5 x^4) / 2
  kq = fmaf(kq, x2, 0.39873274f);
  const float dens = fmaf(-t, t, 1.f) * kq;             // Phi'(x); 0 in both tails, where the clamped k' no longer matters
  return fmaf(x, dens, fmaf(0.5f, t, 0.5f));
}

// ---- MLP dropout (models/vit.py:122,125-139: nn.Dropout after the activation and after fc2) -----------------------------
// Counter-based keep-mask, regenerated in backward instead of stored: element (row, col) of an [M, N] tensor belongs to
// group g = row * ceil(N / 4) + col / 4; one splitmix64 of key + g * golden gives four 16-bit uniforms, lane col % 4 is
// kept iff it is >= thr = round(p * 65536) and the kept value is scaled by 1 / (1 - thr / 65536).  key = *seed + offset:
// the seed lives in device memory so that a captured CUDA graph draws a fresh mask on every replay, the offset tells
// layers and dropout sites apart.  oracle/mhla_oracle.py:mlp_dropout_keep_mask reproduces it bit for bit.
struct DropSpec {
  const unsigned long long* seed = nullptr;  // device pointer; nullptr = no dropout
  unsigned long long offset = 0;
  unsigned thr = 0;       // drop iff u16 < thr
  float inv_keep = 1.f;
  int groups_per_row = 0;
};
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// v[0..31] = columns col..col+31 (col % 4 == 0) of row `row`
__device__ __forceinline__ void dropout32(float (&v)[32], unsigned long long key, int row, int col, const DropSpec& d) {
  const unsigned long long g0 = (unsigned long long)row * (unsigned long long)d.groups_per_row + (unsigned)(col >> 2);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const unsigned long long z = splitmix64(key + (g0 + j) * 0x9E3779B97F4A7C15ull);
    const unsigned lo = (unsigned)z, hi = (unsigned)(z >> 32);
    v[4 * j] = (lo & 0xffffu) >= d.thr ? v[4 * j] * d.inv_keep : 0.f;
    v[4 * j + 1] = (lo >> 16) >= d.thr ? v[4 * j + 1] * d.inv_keep : 0.f;
    v[4 * j + 2] = (hi & 0xffffu) >= d.thr ? v[4 * j + 2] * d.inv_keep : 0.f;
    v[4 * j + 3] = (hi >> 16) >= d.thr ? v[4 * j + 3] * d.inv_keep : 0.f;
  }
}

}  // namespace favit
