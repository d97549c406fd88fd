// Patch-embedding + superpixel pooling, algebraically fused.
//
// /root/reference/models/sppp_mhla.py:281-300 projects every patch (models/vit.py:36-41: einops
// 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' + Linear) and then averages the embeddings of the patches of each superpixel
// (models/sppp.py:209-210).  Both steps are linear, so   mean_p (W x_p + b) = W (mean_p x_p) + b :
// this kernel averages the RAW pixel patches of every superpixel straight out of the image, and the projection then
// runs on R rows per image instead of P (12x fewer at 224 px / patch 16 / 16 superpixels, 64x at 512 px / patch 8 / 64).
// The [B, P, D] embedding tensor, its pooling backward and the P-row weight-gradient GEMM never exist.
//
// One CTA per (image, slot).  Threads walk the 3 ps^2 features of a pixel patch in MEMORY order (channel, row, column:
// ps consecutive floats = 64 bytes at ps 16, whole 32-byte sectors at ps 8) and add the slot's patches in ascending
// patch order — the CSR of favit_sppp_assign — so the result is deterministic; the feature vector is transposed to the
// Linear's (p1 p2 c) order in shared memory and stored coalesced.  HBM-bound: every image byte is read exactly once.
#include "favit_common.cuh"

namespace favit {
namespace {

// VEC = 4: a thread owns four consecutive pixels of a patch row (one 16-byte load per patch; needs patch % 4 == 0,
// img_w % 4 == 0 and a 16-byte aligned image), VEC = 1: one pixel (any geometry).
template <typename TOut, int VEC>
__global__ void __launch_bounds__(256) sppp_pool_pixels_kernel(const float* __restrict__ img,
                                                               const int32_t* __restrict__ order,
                                                               const int32_t* __restrict__ offsets,
                                                               const int32_t* __restrict__ num_slots,
                                                               TOut* __restrict__ out, int C, int img_h, int img_w, int ps,
                                                               int grid, int P, int R, int r_cap) {
  extern __shared__ float s_dyn[];
  const int F = ps * ps * C;
  float* s_out = s_dyn;                                  // [F] in (p1 p2 c) order
  int* s_off = reinterpret_cast<int*>(s_dyn + F);        // [patches of this slot] pixel offset of each patch's corner
  const int b = blockIdx.x / R, r = blockIdx.x - b * R;
  const int ns = min(min(num_slots[b], r_cap), R);
  const int beg = r < ns ? offsets[(int64_t)b * (r_cap + 1) + r] : 0;
  const int end = r < ns ? offsets[(int64_t)b * (r_cap + 1) + r + 1] : 0;
  const int n = end - beg;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int p = order[(int64_t)b * P + beg + t];
    const int i = p / grid, j = p - i * grid;
    s_off[t] = i * ps * img_w + j * ps;
  }
  __syncthreads();
  const float inv = 1.f / (float)max(n, 1);
  const int pp = ps * ps;
  // The kernel is a latency-bound gather of short segments (ps consecutive floats per patch row): bytes in flight are
  // what buys bandwidth, so every thread keeps UN independent loads going (UN patches of its own pixel group).  The
  // summation order per feature is fixed: (t mod UN) partial sums, combined in a fixed tree.
  constexpr int UN = 8;
  const int groups = F / VEC;   // pixel groups in memory order (c, p1, p2 / VEC)
  for (int e = threadIdx.x; e < groups; e += blockDim.x) {
    const int q = e * VEC;
    const int c = q / pp, rem = q - c * pp;
    const int p1 = rem / ps, p2 = rem - p1 * ps;
    const float* base = img + ((int64_t)(b * C + c) * img_h + p1) * img_w + p2;
    float a[UN][VEC];
#pragma unroll
    for (int k = 0; k < UN; ++k)
#pragma unroll
      for (int v = 0; v < VEC; ++v) a[k][v] = 0.f;
    int t = 0;
    for (; t + UN <= n; t += UN) {
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        if constexpr (VEC == 4) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(base + s_off[t + k]));
          a[k][0] += x.x; a[k][1] += x.y; a[k][2] += x.z; a[k][3] += x.w;
        } else {
          a[k][0] += __ldg(base + s_off[t + k]);
        }
      }
    }
    for (; t < n; ++t) {
      if constexpr (VEC == 4) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(base + s_off[t]));
        a[0][0] += x.x; a[0][1] += x.y; a[0][2] += x.z; a[0][3] += x.w;
      } else {
        a[0][0] += __ldg(base + s_off[t]);
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float sum = ((a[0][v] + a[1][v]) + (a[2][v] + a[3][v])) + ((a[4][v] + a[5][v]) + (a[6][v] + a[7][v]));
      s_out[(p1 * ps + p2 + v) * C + c] = sum * inv;
    }
  }
  __syncthreads();
  TOut* o = out + (int64_t)blockIdx.x * F;
  for (int f = threadIdx.x; f < F; f += blockDim.x) Elem<TOut>::st(o + f, s_out[f]);
}

}  // namespace
}  // namespace favit

using namespace favit;

extern "C" int favit_sppp_pool_pixels(const float* image, int B, int C, int img_h, int img_w, int patch, int grid,
                                      const int32_t* order, const int32_t* offsets, const int32_t* num_slots, void* out,
                                      favit_dtype out_dtype, int R, int r_cap, favit_stream stream) {
  FAVIT_CHECK_ARG(image && order && offsets && num_slots && out, "sppp_pool_pixels: null pointer");
  FAVIT_CHECK_ARG(B > 0 && C > 0 && patch > 0 && grid > 0 && R > 0 && r_cap > 0, "sppp_pool_pixels: sizes must be > 0");
  FAVIT_CHECK_ARG((int64_t)grid * patch <= img_h && (int64_t)grid * patch <= img_w,
                  "sppp_pool_pixels: grid*patch (%d) exceeds the image (%dx%d)", grid * patch, img_h, img_w);
  FAVIT_CHECK_ARG((int64_t)B * R < INT32_MAX && (int64_t)B * C * img_h * img_w < ((int64_t)1 << 40),
                  "sppp_pool_pixels: problem too large");
  const int P = grid * grid;
  const size_t smem = ((size_t)patch * patch * C + (size_t)P) * 4;
  if (smem > 200 * 1024) {
    set_error("sppp_pool_pixels: %zu bytes of shared memory per CTA (patch %d, %d channels, %d patches) unsupported", smem,
              patch, C, P);
    return FAVIT_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (patch % 4 == 0) && (img_w % 4 == 0) && ((uintptr_t)image % 16 == 0);
  static bool configured = false;
  if (!configured) {
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(sppp_pool_pixels_kernel<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(sppp_pool_pixels_kernel<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(sppp_pool_pixels_kernel<__nv_bfloat16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(sppp_pool_pixels_kernel<__nv_bfloat16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  // one pixel group per thread: narrow pixel patches get narrower CTAs instead of idle lanes
  int threads = (patch * patch * C / (vec ? 4 : 1) + 31) / 32 * 32;
  threads = threads < 32 ? 32 : (threads > 256 ? 256 : threads);
  note_kernel("sppp_pool_pixels_kernel<VEC=%d> grid=%d threads=%d", vec ? 4 : 1, B * R, threads);
#define FAVIT_PIX(T, V)                                                                                              \
  sppp_pool_pixels_kernel<T, V><<<(unsigned)(B * R), threads, smem, st>>>(image, order, offsets, num_slots, (T*)out, C, \
                                                                         img_h, img_w, patch, grid, P, R, r_cap)
  if (out_dtype == FAVIT_BF16) {
    if (vec) FAVIT_PIX(__nv_bfloat16, 4); else FAVIT_PIX(__nv_bfloat16, 1);
  } else if (out_dtype == FAVIT_F32) {
    if (vec) FAVIT_PIX(float, 4); else FAVIT_PIX(float, 1);
  } else { set_error("sppp_pool_pixels: bad dtype"); return FAVIT_ERR_ARG; }
#undef FAVIT_PIX
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

// ---- class token + dynamic positional encoding ------------------------------------------------------------------------
// /root/reference/models/sppp_mhla.py:302-310 and models/sppp.py:271-299: x = cat(cls, pooled); centroids get a (0.5, 0.5)
// row for the class token; pe = cat(sin(cx * freq), cos(cy * freq)), freq_i = exp(-i ln(10000) / (D/2)); x += pe.
// One pass instead of eight torch launches (cat, full, arange, exp, mul, sin, cos, cat, add).  out [B, R+1, D] fp32.
namespace favit {
namespace {
__global__ void __launch_bounds__(256) sppp_embed_tokens_kernel(const float* __restrict__ pooled, const float* __restrict__ cls,
                                                                const float* __restrict__ centroids, float* __restrict__ out,
                                                                int B, int R, int D) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * (R + 1) * D;
  if (idx >= total) return;
  const int d = (int)(idx % D);
  const int t = (int)((idx / D) % (R + 1));
  const int b = (int)(idx / ((int64_t)D * (R + 1)));
  const int half = D / 2;
  float cx = 0.5f, cy = 0.5f, v;
  if (t == 0) {
    v = cls[d];
  } else {
    v = pooled[((int64_t)b * R + (t - 1)) * D + d];
    cx = centroids[((int64_t)b * R + (t - 1)) * 2];
    cy = centroids[((int64_t)b * R + (t - 1)) * 2 + 1];
  }
  const int i = d < half ? d : d - half;
  const float freq = expf((float)i * (-9.210340371976184f / (float)half));
  const float pe = d < half ? sinf(cx * freq) : cosf(cy * freq);
  out[idx] = v + pe;
}
}  // namespace
}  // namespace favit

extern "C" int favit_sppp_embed_tokens(const float* pooled, const float* cls_token, const float* centroids, float* out, int B,
                                       int R, int D, favit_stream stream) {
  FAVIT_CHECK_ARG(pooled && cls_token && centroids && out, "sppp_embed_tokens: null pointer");
  FAVIT_CHECK_ARG(B > 0 && R > 0 && D > 0 && D % 2 == 0, "sppp_embed_tokens: sizes must be > 0 and D even");
  const int64_t total = (int64_t)B * (R + 1) * D;
  sppp_embed_tokens_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, (cudaStream_t)stream>>>(pooled, cls_token, centroids, out,
                                                                                           B, R, D);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

// ---- patchify: 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (models/vit.py:38-39) + cast to the GEMM operand type ---------------
// One pass: fp32 image in, [B*P, p*p*C] rows in the compute dtype out (torch: a permuted fp32 copy, then a cast).
// Thread -> one output element group of 4 consecutive p2 (so reads are 16-byte, coalesced along image rows); the
// transposition to the feature order happens through the index arithmetic, writes are 2- / 4-byte strided by C.
namespace favit {
namespace {
template <typename TOut>
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, TOut* __restrict__ out, int B, int C,
                                                       int S, int ps, int g) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B * C * S * S / 4
  const int64_t total = (int64_t)B * C * S * S / 4;
  if (idx >= total) return;
  const int x4 = (int)(idx % (S / 4)) * 4;
  const int y = (int)((idx / (S / 4)) % S);
  const int c = (int)((idx / ((int64_t)(S / 4) * S)) % C);
  const int b = (int)(idx / ((int64_t)(S / 4) * S * C));
  const float4 v = *reinterpret_cast<const float4*>(img + (((int64_t)b * C + c) * S + y) * S + x4);
  const int pi = y / ps, p1 = y - pi * ps;
  const int pj = x4 / ps, p2 = x4 - pj * ps;   // ps % 4 == 0: the four pixels stay inside one patch
  const int F = ps * ps * C;
  TOut* o = out + ((int64_t)b * g * g + (int64_t)pi * g + pj) * F + (int64_t)(p1 * ps + p2) * C + c;
  Elem<TOut>::st(o, v.x);
  Elem<TOut>::st(o + C, v.y);
  Elem<TOut>::st(o + 2 * C, v.z);
  Elem<TOut>::st(o + 3 * C, v.w);
}

// C == 3 (every model of the reference): a thread takes four pixels of ALL THREE channel planes (three 16-byte loads) and
// writes their twelve interleaved features as one contiguous run (24 bytes in bf16, 48 in fp32); consecutive threads write
// consecutive runs, so both sides of the transposition are coalesced.
template <typename TOut>
__global__ void __launch_bounds__(256) patchify3_kernel(const float* __restrict__ img, TOut* __restrict__ out, int B, int S,
                                                        int ps, int g) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over B * S * S / 4
  const int64_t total = (int64_t)B * S * S / 4;
  if (idx >= total) return;
  const int x4 = (int)(idx % (S / 4)) * 4;
  const int y = (int)((idx / (S / 4)) % S);
  const int b = (int)(idx / ((int64_t)(S / 4) * S));
  const int64_t plane = (int64_t)S * S;
  const float* src = img + (int64_t)b * 3 * plane + (int64_t)y * S + x4;
  const float4 r = __ldcs(reinterpret_cast<const float4*>(src));
  const float4 gch = __ldcs(reinterpret_cast<const float4*>(src + plane));
  const float4 bl = __ldcs(reinterpret_cast<const float4*>(src + 2 * plane));
  const int pi = y / ps, p1 = y - pi * ps;
  const int pj = x4 / ps, p2 = x4 - pj * ps;
  const int F = ps * ps * 3;
  TOut* o = out + ((int64_t)b * g * g + (int64_t)pi * g + pj) * F + (int64_t)(p1 * ps + p2) * 3;
  const float f[12] = {r.x, gch.x, bl.x, r.y, gch.y, bl.y, r.z, gch.z, bl.z, r.w, gch.w, bl.w};
  if constexpr (sizeof(TOut) == 2) {   // 24 bytes, 8-byte aligned (F and 3 * p2 are multiples of 4 elements)
    uint2* o2 = reinterpret_cast<uint2*>(o);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      o2[i] = make_uint2(pack_bf16x2(f[4 * i], f[4 * i + 1]), pack_bf16x2(f[4 * i + 2], f[4 * i + 3]));
  } else {
    float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
    for (int i = 0; i < 3; ++i) o4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  }
}
}  // namespace
}  // namespace favit

extern "C" int favit_patchify(const float* image, void* out, favit_dtype out_dtype, int B, int C, int S, int patch,
                              favit_stream stream) {
  FAVIT_CHECK_ARG(image && out && B > 0 && C > 0 && S > 0 && patch > 0, "patchify: bad argument");
  FAVIT_CHECK_ARG(S % patch == 0 && patch % 4 == 0 && ((uintptr_t)image % 16 == 0),
                  "patchify: needs S %% patch == 0, patch %% 4 == 0 and a 16-byte aligned image (S=%d patch=%d)", S, patch);
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 3 && ((uintptr_t)out % 16 == 0) && (out_dtype == FAVIT_BF16 || out_dtype == FAVIT_F32)) {
    const unsigned blocks3 = (unsigned)ceil_div64((int64_t)B * S * S / 4, 256);
    if (out_dtype == FAVIT_BF16)
      patchify3_kernel<__nv_bfloat16><<<blocks3, 256, 0, st>>>(image, (__nv_bfloat16*)out, B, S, patch, S / patch);
    else
      patchify3_kernel<float><<<blocks3, 256, 0, st>>>(image, (float*)out, B, S, patch, S / patch);
    FAVIT_CHECK_LAUNCH();
    return FAVIT_OK;
  }
  const int64_t total = (int64_t)B * C * S * S / 4;
  const unsigned blocks = (unsigned)ceil_div64(total, 256);
  if (out_dtype == FAVIT_BF16)
    patchify_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(image, (__nv_bfloat16*)out, B, C, S, patch, S / patch);
  else if (out_dtype == FAVIT_F32)
    patchify_kernel<float><<<blocks, 256, 0, st>>>(image, (float*)out, B, C, S, patch, S / patch);
  else { set_error("patchify: bad dtype"); return FAVIT_ERR_ARG; }
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}
