// Shared helpers for the favit_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/favit.h"

namespace favit {

// thread-local last-error message (favit_last_error)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// which kernel variant the dispatching entry point chose (favit_last_kernel; tests assert on it)
void note_kernel(const char* fmt, ...);
// make the primary context current on this thread (once per thread) before driver-API calls
void bind_context();

#define FAVIT_CHECK_ARG(cond, ...)                    \
  do {                                                \
    if (!(cond)) {                                    \
      ::favit::set_error(__VA_ARGS__);                \
      return FAVIT_ERR_ARG;                           \
    }                                                 \
  } while (0)

#define FAVIT_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::favit::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                         __LINE__);                                                         \
      return FAVIT_ERR_CUDA;                                                                \
    }                                                                                       \
  } while (0)

#define FAVIT_CHECK_LAUNCH()                                                                \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError(); /* clears it: one bad launch must not poison the rest */ \
    if (_e != cudaSuccess) {                                                                \
      ::favit::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),        \
                         __FILE__, __LINE__);                                               \
      return FAVIT_ERR_CUDA;                                                                \
    }                                                                                       \
    ::favit::count_launch();                                                                \
  } while (0)

int num_sms();

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// A kernel launched with the programmatic-stream-serialization attribute may become resident while its predecessor in the
// stream is still running; it must call pdl_wait() before its FIRST global-memory access (reads of what the predecessor
// wrote, and writes the predecessor may still be reading), and calls pdl_launch_dependents() early so that ITS successor
// can do the same.  For the 10-15 us SPPP kernels this hides the CTA ramp and the launch gap behind the predecessor's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- element access helpers -------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 8 consecutive elements <-> 8 floats (16-byte vector for bf16, 2x16-byte for fp32). Pointers must
// be 16-byte aligned.
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

}  // namespace favit
