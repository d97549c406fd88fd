// GPU superpixel segmentation (SLIC) — the step upstream of the hot path.
//
// /root/reference/models/sppp.py:26-74 calls skimage.segmentation.slic per image on the CPU (device -> host copy,
// Python loop, host -> device copy of an int64 map).  This is the same algorithm family on the device, for the whole
// batch at once, written so that a CPU restatement (oracle/slic_oracle.py) reproduces it BIT FOR BIT:
//   1. Gaussian pre-smoothing (sigma, truncated at 4 sigma, 'nearest' borders), separable, explicit mul/add order
//   2. K' = gy x gx cluster centres on a regular grid, colour = the smoothed pixel at the centre
//   3. `iters` rounds of { assign every pixel to the nearest of the 3 x 3 grid-neighbour centres, distance =
//      sum_c (f_c - mu_c)^2 / compactness^2 + ((y - cy)^2 + (x - cx)^2) / step^2, ties -> lowest cluster index;
//      recompute the centres } with 64-bit fixed-point sums (exact, order independent -> deterministic)
//   4. labels int64 [B, H, W] in [0, K')
// Deviations from scikit-image's slic (which is not installed here, so nothing can be pinned against it): no RGB -> Lab
// conversion (the reference feeds mean / std normalised tensors, outside Lab's [0, 1] domain), the search window is the
// 3 x 3 grid neighbourhood of the pixel instead of +-2 steps around each centre, no connectivity enforcement pass.
#include "favit_common.cuh"

namespace favit {
namespace {

constexpr int kMaxRadius = 8;
constexpr float kFix = 65536.f;   // fixed-point scale of the centre sums

struct BlurTaps {
  float w[2 * kMaxRadius + 1];
  int radius;
};

// one pass of the separable blur along x (DIR = 0) or y (DIR = 1); taps applied in ascending offset order
template <int DIR>
__global__ void __launch_bounds__(256) slic_blur_kernel(const float* __restrict__ in, float* __restrict__ out, int planes,
                                                        int H, int W, const BlurTaps taps) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)planes * H * W) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H);
  const int64_t plane = idx / ((int64_t)H * W);
  const float* src = in + plane * H * W;
  float acc = 0.f;
  for (int t = -taps.radius; t <= taps.radius; ++t) {
    const int yy = DIR == 1 ? min(max(y + t, 0), H - 1) : y;
    const int xx = DIR == 0 ? min(max(x + t, 0), W - 1) : x;
    acc = __fadd_rn(acc, __fmul_rn(taps.w[t + taps.radius], src[(int64_t)yy * W + xx]));
  }
  out[idx] = acc;
}

// centres: [B][K'][2 + C] = (cy, cx, colour...)
__global__ void slic_init_kernel(const float* __restrict__ feat, float* __restrict__ centres, int B, int C, int H, int W,
                                 int gy, int gx, float step_y, float step_x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int K = gy * gx;
  if (i >= B * K) return;
  const int b = i / K, k = i - b * K;
  const int iy = k / gx, ix = k - iy * gx;
  const float cy = __fmul_rn((float)iy + 0.5f, step_y), cx = __fmul_rn((float)ix + 0.5f, step_x);
  const int py = min((int)cy, H - 1), px = min((int)cx, W - 1);
  float* c = centres + (int64_t)i * (2 + C);
  c[0] = cy;
  c[1] = cx;
  for (int ch = 0; ch < C; ++ch) c[2 + ch] = feat[((int64_t)(b * C + ch) * H + py) * W + px];
}

template <int C>
__global__ void __launch_bounds__(256) slic_assign_kernel(const float* __restrict__ feat, const float* __restrict__ centres,
                                                          int64_t* __restrict__ labels, long long* __restrict__ sums,
                                                          int B, int H, int W, int gy, int gx, float inv_step_y,
                                                          float inv_step_x, float inv_step2, float inv_comp2,
                                                          int accumulate) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * H * W) return;
  const int x = (int)(idx % W), y = (int)((idx / W) % H), b = (int)(idx / ((int64_t)H * W));
  const int K = gy * gx;
  float f[C];
#pragma unroll
  for (int ch = 0; ch < C; ++ch) f[ch] = feat[((int64_t)(b * C + ch) * H + y) * W + x];
  const int cyi = min((int)__fmul_rn((float)y, inv_step_y), gy - 1), cxi = min((int)__fmul_rn((float)x, inv_step_x), gx - 1);
  float best = 3.402823466e38f;
  int bestk = 0;
  for (int dy = -1; dy <= 1; ++dy) {
    const int iy = cyi + dy;
    if (iy < 0 || iy >= gy) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int ix = cxi + dx;
      if (ix < 0 || ix >= gx) continue;
      const int k = iy * gx + ix;
      const float* c = centres + ((int64_t)b * K + k) * (2 + C);
      float dc = 0.f;
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        const float d = __fsub_rn(f[ch], c[2 + ch]);
        dc = __fadd_rn(dc, __fmul_rn(d, d));
      }
      const float ey = __fsub_rn((float)y, c[0]), ex = __fsub_rn((float)x, c[1]);
      const float ds = __fadd_rn(__fmul_rn(ey, ey), __fmul_rn(ex, ex));
      const float dist = __fadd_rn(__fmul_rn(dc, inv_comp2), __fmul_rn(ds, inv_step2));
      if (dist < best) {   // candidates are visited in ascending k: ties keep the lowest index
        best = dist;
        bestk = k;
      }
    }
  }
  labels[idx] = bestk;
  if (accumulate) {
    long long* s = sums + ((int64_t)b * K + bestk) * (3 + C);
    atomicAdd((unsigned long long*)s, 1ull);
    atomicAdd((unsigned long long*)(s + 1), (unsigned long long)y);
    atomicAdd((unsigned long long*)(s + 2), (unsigned long long)x);
#pragma unroll
    for (int ch = 0; ch < C; ++ch)
      atomicAdd((unsigned long long*)(s + 3 + ch), (unsigned long long)(long long)llrintf(__fmul_rn(f[ch], kFix)));
  }
}

__global__ void slic_update_kernel(const long long* __restrict__ sums, float* __restrict__ centres, int total, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long* s = sums + (int64_t)i * (3 + C);
  const long long n = s[0];
  if (n <= 0) return;   // an empty cluster keeps its centre
  float* c = centres + (int64_t)i * (2 + C);
  c[0] = (float)((double)s[1] / (double)n);
  c[1] = (float)((double)s[2] / (double)n);
  for (int ch = 0; ch < C; ++ch) c[2 + ch] = (float)((double)s[3 + ch] / (double)n / (double)kFix);
}

}  // namespace
}  // namespace favit

using namespace favit;

extern "C" int favit_slic_grid(int H, int W, int n_segments, int* gy, int* gx) {
  FAVIT_CHECK_ARG(H > 0 && W > 0 && n_segments > 0 && gy && gx, "slic_grid: bad argument");
  int x = (int)lrint(sqrt((double)n_segments * (double)W / (double)H));
  x = x < 1 ? 1 : (x > W ? W : x);
  int y = (int)lrint((double)n_segments / (double)x);
  y = y < 1 ? 1 : (y > H ? H : y);
  *gy = y;
  *gx = x;
  return FAVIT_OK;
}

extern "C" int favit_slic_segment(const float* image, int B, int C, int H, int W, int n_segments, float compactness,
                                  float sigma, int iters, int64_t* labels, float* feat, float* tmp, float* centres,
                                  long long* sums, favit_stream stream) {
  FAVIT_CHECK_ARG(image && labels && feat && tmp && centres && sums, "slic_segment: null pointer");
  FAVIT_CHECK_ARG(B > 0 && H > 0 && W > 0 && n_segments > 0 && compactness > 0.f && sigma >= 0.f && iters >= 0,
                  "slic_segment: bad sizes / parameters");
  FAVIT_CHECK_ARG(C == 1 || C == 3, "slic_segment: 1 or 3 channels (got %d)", C);
  FAVIT_CHECK_ARG((int64_t)B * C * H * W < ((int64_t)1 << 40), "slic_segment: problem too large");
  cudaStream_t st = (cudaStream_t)stream;
  int gy, gx;
  favit_slic_grid(H, W, n_segments, &gy, &gx);
  const int K = gy * gx;
  const float step_y = (float)H / (float)gy, step_x = (float)W / (float)gx;
  const float step = step_y > step_x ? step_y : step_x;
  const int64_t npix = (int64_t)B * H * W, nval = npix * C;
  const unsigned gb = (unsigned)ceil_div64(nval, 256);
  // 1. smoothing
  BlurTaps taps;
  int radius = (int)(4.0 * (double)sigma + 0.5);
  radius = radius > kMaxRadius ? kMaxRadius : radius;
  taps.radius = radius;
  if (sigma > 0.f && radius > 0) {
    double w[2 * kMaxRadius + 1], tot = 0.0;
    for (int t = -radius; t <= radius; ++t) { w[t + radius] = exp(-0.5 * (double)t * t / ((double)sigma * sigma)); tot += w[t + radius]; }
    for (int t = 0; t <= 2 * radius; ++t) taps.w[t] = (float)(w[t] / tot);
    slic_blur_kernel<0><<<gb, 256, 0, st>>>(image, tmp, B * C, H, W, taps);
    FAVIT_CHECK_LAUNCH();
    slic_blur_kernel<1><<<gb, 256, 0, st>>>(tmp, feat, B * C, H, W, taps);
    FAVIT_CHECK_LAUNCH();
  } else {
    FAVIT_CHECK_CUDA(cudaMemcpyAsync(feat, image, (size_t)nval * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  // 2. centres
  slic_init_kernel<<<(unsigned)ceil_div(B * K, 128), 128, 0, st>>>(feat, centres, B, C, H, W, gy, gx, step_y, step_x);
  FAVIT_CHECK_LAUNCH();
  // 3. iterations (the last assignment only labels)
  const float inv_step2 = 1.f / (step * step), inv_comp2 = 1.f / (compactness * compactness);
  const unsigned pb = (unsigned)ceil_div64(npix, 256);
  for (int it = 0; it <= iters; ++it) {
    const int acc = it < iters ? 1 : 0;
    if (acc) FAVIT_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * K * (3 + C) * sizeof(long long), st));
    if (C == 3)
      slic_assign_kernel<3><<<pb, 256, 0, st>>>(feat, centres, labels, sums, B, H, W, gy, gx, 1.f / step_y, 1.f / step_x,
                                                 inv_step2, inv_comp2, acc);
    else
      slic_assign_kernel<1><<<pb, 256, 0, st>>>(feat, centres, labels, sums, B, H, W, gy, gx, 1.f / step_y, 1.f / step_x,
                                                 inv_step2, inv_comp2, acc);
    FAVIT_CHECK_LAUNCH();
    if (acc) {
      slic_update_kernel<<<(unsigned)ceil_div(B * K, 128), 128, 0, st>>>(sums, centres, B * K, C);
      FAVIT_CHECK_LAUNCH();
    }
  }
  return FAVIT_OK;
}
