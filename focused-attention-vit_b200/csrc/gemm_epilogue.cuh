// Element-wise epilogue math shared by the SIMT and tcgen05 GEMMs.
#pragma once
#include <cuda_runtime.h>

namespace favit {

// nn.GELU() default (erf form), reference models/vit.py:120, models/mhla.py:199.
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

}  // namespace favit
