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

// bf16 path: erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7 before the two MUFU approximations, about 1e-6 after:
// three orders of magnitude below bf16 resolution) — 1 MUFU.RCP + 1 MUFU.EX2 + 9 FMA-class instructions instead of
// erff's ~25, which matters because the GELU epilogue must drain a 128 x 256 tile faster than the MMAs fill the next.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// returns erf(|z|) and exp(-z^2) for z = x / sqrt(2)
__device__ __forceinline__ void erf_exp_fast(float x, float& erf_abs, float& e) {
  const float az = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, az, 1.f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  e = ex2_approx(az * az * -1.4426950408889634f);
  erf_abs = fmaf(-p, e, 1.f);
}
__device__ __forceinline__ float gelu_fast(float x) {
  float er, e;
  erf_exp_fast(x, er, e);
  return 0.5f * x * (1.f + copysignf(er, x));
}
__device__ __forceinline__ float dgelu_fast(float x) {
  float er, e;
  erf_exp_fast(x, er, e);
  return fmaf(x * 0.3989422804014327f, e, 0.5f * (1.f + copysignf(er, x)));
}

}  // namespace favit
