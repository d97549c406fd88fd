// Multi-tensor AdamW — the optimizer step of the training loop the hot path sits in
// (/root/reference/experiments/mhla_pretrained.py:320-327, 367: torch.optim.AdamW over THREE parameter groups,
// `latent_proj` at 5x the learning rate; main.py:129-132: lr 1e-4, weight decay 0.05).
//
// One launch updates up to 40 tensors, each with its own learning rate and weight decay (so parameter groups cost
// nothing), reads the step count from device memory (a captured CUDA graph advances it in-stream) and can fold a
// gradient scale in (1 / world size when the gradient all-reduce sums instead of averaging): p, m, v are read and
// written once, g is read once — 28 bytes per parameter, the floor for fp32 AdamW state.
//   p <- p (1 - lr wd);  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2
//   p <- p - lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)          (torch.optim.AdamW, decoupled decay)
#include "favit_common.cuh"

namespace favit {
namespace {

constexpr int kAdamMax = 40;
struct AdamTable {
  float* p[kAdamMax];
  const float* g[kAdamMax];
  float* m[kAdamMax];
  float* v[kAdamMax];
  long long n[kAdamMax];
  float lr[kAdamMax];
  float wd[kAdamMax];
};

__global__ void __launch_bounds__(256) adamw_multi_kernel(const __grid_constant__ AdamTable t,
                                                          const long long* __restrict__ step, double beta1_d, double beta2_d,
                                                          float eps, float grad_scale) {
  const int ti = blockIdx.y;
  const long long n = t.n[ti];
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  // hyper-parameter arithmetic in double, like the Python side of torch.optim.AdamW: 1 - 0.999f is off by 5e-5 in fp32
  const double s = (double)__ldg(step);
  const float bc1 = (float)(1.0 - pow(beta1_d, s)), bc2 = (float)(1.0 - pow(beta2_d, s));
  const float beta1 = (float)beta1_d, beta2 = (float)beta2_d;
  const float omb1 = (float)(1.0 - beta1_d), omb2 = (float)(1.0 - beta2_d);
  const float lr = t.lr[ti], decay = 1.f - lr * t.wd[ti];
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  float* __restrict__ p = t.p[ti];
  const float* __restrict__ g = t.g[ti];
  float* __restrict__ m = t.m[ti];
  float* __restrict__ v = t.v[ti];
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0;
  for (; i < n; i += stride) {
    float pp[4], gg[4], mm[4], vv[4];
    const int k = (int)min((long long)4, n - i);
    if (vec && k == 4) {
      *reinterpret_cast<float4*>(pp) = *reinterpret_cast<const float4*>(p + i);
      *reinterpret_cast<float4*>(gg) = __ldcs(reinterpret_cast<const float4*>(g + i));
      *reinterpret_cast<float4*>(mm) = *reinterpret_cast<const float4*>(m + i);
      *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + i);
    } else {
      for (int e = 0; e < 4; ++e) {
        pp[e] = e < k ? p[i + e] : 0.f; gg[e] = e < k ? g[i + e] : 0.f;
        mm[e] = e < k ? m[i + e] : 0.f; vv[e] = e < k ? v[i + e] : 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ge = gg[e] * grad_scale;
      pp[e] *= decay;
      mm[e] = beta1 * mm[e] + omb1 * ge;
      vv[e] = beta2 * vv[e] + omb2 * ge * ge;
      const float denom = sqrtf(vv[e]) * inv_sqrt_bc2 + eps;
      pp[e] -= step_size * (mm[e] / denom);
    }
    if (vec && k == 4) {
      *reinterpret_cast<float4*>(p + i) = *reinterpret_cast<float4*>(pp);
      *reinterpret_cast<float4*>(m + i) = *reinterpret_cast<float4*>(mm);
      *reinterpret_cast<float4*>(v + i) = *reinterpret_cast<float4*>(vv);
    } else {
      for (int e = 0; e < k; ++e) { p[i + e] = pp[e]; m[i + e] = mm[e]; v[i + e] = vv[e]; }
    }
  }
}

}  // namespace
}  // namespace favit

using namespace favit;

extern "C" int favit_adamw_multi(int count, void* const* params, const void* const* grads, void* const* exp_avg,
                                 void* const* exp_avg_sq, const int64_t* numel, const float* lr, const float* weight_decay,
                                 const int64_t* step, double beta1, double beta2, float eps, float grad_scale,
                                 favit_stream stream) {
  FAVIT_CHECK_ARG(count > 0 && params && grads && exp_avg && exp_avg_sq && numel && lr && weight_decay && step,
                  "adamw_multi: bad argument");
  FAVIT_CHECK_ARG(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0.f, "adamw_multi: bad hyper-parameters");
  cudaStream_t st = (cudaStream_t)stream;
  for (int t0 = 0; t0 < count; t0 += kAdamMax) {
    const int n = count - t0 < kAdamMax ? count - t0 : kAdamMax;
    AdamTable t;
    long long big = 0;
    for (int i = 0; i < kAdamMax; ++i) {
      const int k = i < n ? t0 + i : t0;
      FAVIT_CHECK_ARG(params[k] && grads[k] && exp_avg[k] && exp_avg_sq[k] && numel[k] >= 0, "adamw_multi: null tensor %d", k);
      t.p[i] = (float*)params[k];
      t.g[i] = (const float*)grads[k];
      t.m[i] = (float*)exp_avg[k];
      t.v[i] = (float*)exp_avg_sq[k];
      t.n[i] = i < n ? numel[k] : 0;
      t.lr[i] = lr[k];
      t.wd[i] = weight_decay[k];
      big = t.n[i] > big ? t.n[i] : big;
    }
    long long bx = (big + 4095) / 4096;   // 4 elements x 256 threads x 4 iterations per CTA
    bx = bx < 1 ? 1 : (bx > 2 * num_sms() ? 2 * num_sms() : bx);
    adamw_multi_kernel<<<dim3((unsigned)bx, n), 256, 0, st>>>(t, (const long long*)step, beta1, beta2, eps, grad_scale);
    FAVIT_CHECK_LAUNCH();
  }
  return FAVIT_OK;
}
