// bf16 GEMM on the 5th-generation tensor cores of sm_100a: tcgen05.mma with TMEM accumulators, operands
// staged in shared memory by TMA, mbarrier pipelines, one persistent CTA per SM.
//
// It carries the dense contractions of the MHLA block — /root/reference/models/mhla.py:100 (qkv) and :158
// (proj), with the latent projection :105-106 folded into their weights by the host — forward, dgrad and
// wgrad, and (the "next" row) the MLP of models/vit.py:107-139.
//
//   C[M,N] (+)= sum_k A(m,k) * B(n,k)          fp32 accumulation in TMEM
//
// Operand storage (both bf16, row-major tensors in global memory):
//   a_mn == 0 : A stored [M rows][K cols]  ("K-major",  the activation of a forward linear)
//   a_mn == 1 : A stored [K rows][M cols]  ("MN-major", dY in a weight gradient)
//   b_mn == 0 : B stored [N rows][K cols]  (nn.Linear weight [out,in] in a forward linear)
//   b_mn == 1 : B stored [K rows][N cols]  (the weight in a dgrad, the activation in a wgrad)
// Either way a pipeline stage holds a 128 x 64 slab of A and a BN x 64 slab of B in the canonical
// SWIZZLE_128B layout, written by TMA and read by the UMMA shared-memory descriptors.
//
// Warp roles (384 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warp 2 = TMEM
// allocator, warps 4..11 = epilogue: warp w may touch TMEM lanes 32*(w%4).. (32 rows of the tile); the two warps of a
// lane quarter split the tile's columns, so the bias/GELU/residual math drains a tile in half the time.
// TMEM holds two BN-column accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
// Split-K work units (weight gradients: few output tiles, very long reduction) accumulate with fp32
// atomics (red.global.add.f32) into an output the caller has zeroed or wants to accumulate into.
#include <cuda.h>

#include <mutex>

#include "favit_common.cuh"
#include "gemm_epilogue.cuh"
#include "gemm_tcgen05.h"

namespace favit {
namespace tc {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kThreads = 384;        // 4 control warps + 8 epilogue warps
constexpr int kEpiWarps = 8;
constexpr uint32_t kSlabBytes = 64 * 128;  // MN-major: 64 reduction rows x 128 B

__host__ __device__ constexpr int stages_for(int bn) { return bn >= 256 ? 4 : (bn >= 128 ? 6 : 8); }
__host__ __device__ constexpr uint32_t stage_bytes(int bn) { return (uint32_t)(BM * BK * 2 + bn * BK * 2); }
__host__ __device__ constexpr uint32_t smem_bytes_for(int bn) {
  return stages_for(bn) * stage_bytes(bn) + 1024 /*alignment slack*/ + 256 /*barriers*/;
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// relaxed: the hand-over of a TMEM accumulator is ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync; the
// default release form costs a MEMBAR.ALL.CTA per tile per epilogue warp (see tcgen05_ptx.cuh)
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, not as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s at 2 GHz
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (SWIZZLE_128B, sm_100 version field = 1).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, M = 128, N = bn.
__host__ __device__ constexpr uint32_t make_idesc(int bn, int a_mn, int b_mn) {
  return (1u << 4)                     // D format  = F32
         | (1u << 7)                   // A format  = BF16
         | (1u << 10)                  // B format  = BF16
         | ((uint32_t)a_mn << 15)      // A major   (0 = K, 1 = MN)
         | ((uint32_t)b_mn << 16)      // B major
         | ((uint32_t)(bn >> 3) << 17) // N >> 3
         | ((uint32_t)(BM >> 4) << 24);// M >> 4
}

struct KParams {
  int M, N, K;
  int m_tiles, n_tiles, k_blocks, splits, kb_per_split;
  int a_mn, b_mn;
  Epilogue epi;
};

// ---------------------------------------------------------------------------------------------
// epilogue helpers: 32 consecutive columns of one row
// ---------------------------------------------------------------------------------------------
// `vec`: 0 = scalar (ragged / misaligned), 1 = 128-bit accesses, 2 = 256-bit accesses (each lane moves whole 32-byte
// sectors of its own row, so no partially written sector ever reaches L2).
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void load32(const void* base, int dtype, int64_t off, int vec, int ncols, float (&f)[32]) {
  if (dtype == FAVIT_BF16) {
    const __nv_bfloat16* p = (const __nv_bfloat16*)base + off;
    if (vec == 2) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint32_t r[8];
        ldg256(p + 16 * i, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[16 * i + 2 * j] = __uint_as_float(r[j] << 16);
          f[16 * i + 2 * j + 1] = __uint_as_float(r[j] & 0xffff0000u);
        }
      }
    } else if (vec == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float t[8];
        load8(p + 8 * i, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[8 * i + j] = t[j];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = (i < ncols) ? __bfloat162float(p[i]) : 0.f;
    }
  } else {
    const float* p = (const float*)base + off;
    if (vec == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t r[8];
        ldg256(p + 8 * i, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[8 * i + j] = __uint_as_float(r[j]);
      }
    } else if (vec == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(p + 4 * i);
        f[4 * i] = v.x; f[4 * i + 1] = v.y; f[4 * i + 2] = v.z; f[4 * i + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = (i < ncols) ? p[i] : 0.f;
    }
  }
}

__device__ __forceinline__ void store32(void* base, int dtype, int64_t off, int vec, int ncols, const float (&f)[32]) {
  if (dtype == FAVIT_BF16) {
    __nv_bfloat16* p = (__nv_bfloat16*)base + off;
    if (vec == 2) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint32_t r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = pack_bf16x2(f[16 * i + 2 * j], f[16 * i + 2 * j + 1]);
        stg256(p + 16 * i, r);
      }
    } else if (vec == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(f[8 * i], f[8 * i + 1]);
        u.y = pack_bf16x2(f[8 * i + 2], f[8 * i + 3]);
        u.z = pack_bf16x2(f[8 * i + 4], f[8 * i + 5]);
        u.w = pack_bf16x2(f[8 * i + 6], f[8 * i + 7]);
        *reinterpret_cast<uint4*>(p + 8 * i) = u;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < ncols) p[i] = __float2bfloat16_rn(f[i]);
    }
  } else {
    float* p = (float*)base + off;
    if (vec == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(f[8 * i + j]);
        stg256(p + 8 * i, r);
      }
    } else if (vec == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(p + 4 * i) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < ncols) p[i] = f[i];
    }
  }
}

__device__ __forceinline__ int vec_ok(const void* base, int dtype, int64_t ld) {
  const int es = (dtype == FAVIT_BF16) ? 2 : 4;
  if ((((uintptr_t)base) % 32 == 0) && ((ld * es) % 32 == 0)) return 2;
  if ((((uintptr_t)base) % 16 == 0) && ((ld * es) % 16 == 0)) return 1;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const KParams p) {
  constexpr int S = stages_for(BN);
  constexpr uint32_t kABytes = BM * BK * 2;
  constexpr uint32_t kBBytes = BN * BK * 2;
  constexpr uint32_t kStage = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + S * kStage;
  // barrier layout (8 bytes each): full[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM base slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4);
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiWarps);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int total_units = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int tile = u / p.splits, split = u % p.splits;
        const int m0 = (tile / p.n_tiles) * BM, n0 = (tile % p.n_tiles) * BN;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * kStage;
          const uint32_t sb = sa + kABytes;
          mbar_expect_tx(full_bar(stage), kStage);
          const int k0 = kb * BK;
          if (p.a_mn) {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * kSlabBytes, &tmA, full_bar(stage), m0 + c * 64, k0);
          } else {
            tma_load_2d(sa, &tmA, full_bar(stage), k0, m0);
          }
          if (p.b_mn) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * kSlabBytes, &tmB, full_bar(stage), n0 + c * 64, k0);
          } else {
            tma_load_2d(sb, &tmB, full_bar(stage), k0, n0);
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BN, p.a_mn, p.b_mn);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int split = u % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStage;
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 16 k-elements = 32 B along the swizzled 128-B row; 8-row groups 1024 B apart.
            // MN-major: 16 k-rows = 2 KiB; 8-row groups 1024 B apart, 64-wide MN chunks one slab apart.
            const uint64_t da = p.a_mn ? make_smem_desc(sa + k * 2048u, kSlabBytes, 1024u)
                                       : make_smem_desc(sa + k * 32u, 16u, 1024u);
            const uint64_t db = p.b_mn ? make_smem_desc(sb + k * 2048u, kSlabBytes, 1024u)
                                       : make_smem_desc(sb + k * 32u, 16u, 1024u);
            umma_bf16(tmem_c, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem stage once these MMAs have read it
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const Epilogue& e = p.epi;
    const int wq = warp & 3;              // TMEM lane quarter this warp may access
    const int chalf = (warp - 4) >> 2;     // which half of the tile's columns this warp drains
    constexpr int kColsPerWarp = BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int c_vec = vec_ok(e.c, e.c_dtype, e.ldc);
    const int r_vec = e.residual ? vec_ok(e.residual, e.res_dtype, e.ldres) : 0;
    const int x_vec = e.aux ? vec_ok(e.aux, FAVIT_BF16, e.ldaux) : 0;
    const int o_vec = e.aux_out ? vec_ok(e.aux_out, FAVIT_BF16, e.ldaux) : 0;
    const bool b_vec = e.bias ? (((uintptr_t)e.bias) % 16 == 0) : false;
    const bool atomic = (p.splits > 1) || e.accumulate;
    const unsigned long long drop_key = e.drop.seed ? __ldg(e.drop.seed) + e.drop.offset : 0ull;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
      const int tile = u / p.splits;
      const int m0 = (tile / p.n_tiles) * BM, n0 = (tile % p.n_tiles) * BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m0 + wq * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int c0 = chalf * kColsPerWarp; c0 < (chalf + 1) * kColsPerWarp; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_wait_ld();
        const int col = n0 + c0;
        const int ncols = min(32, p.N - col);
        if (row < p.M && ncols > 0) {
          const bool full = ncols == 32;
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (atomic) {
            float* cp = (float*)e.c + (int64_t)row * e.ldc + col;
            if (full && c_vec >= 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp + 4 * i), "f"(v[4 * i]),
                             "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                             : "memory");
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < ncols) atomicAdd(cp + i, v[i]);
            }
          } else {
            if (e.bias) {
              if (full && b_vec) {   // every lane reads the same 128 bytes: broadcast loads
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col) + i);
                  v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < ncols) v[i] += __ldg(e.bias + col + i);
              }
            }
            if (e.act == FAVIT_EPI_GELU) {
              if (e.aux_out) store32(e.aux_out, FAVIT_BF16, (int64_t)row * e.ldaux + col, full ? o_vec : 0, ncols, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
            } else if (e.act == FAVIT_EPI_DGELU_MUL) {
              float x[32];
              load32(e.aux, FAVIT_BF16, (int64_t)row * e.ldaux + col, full ? x_vec : 0, ncols, x);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= dgelu_fast(x[i]);
            }
            if (e.drop.seed) dropout32(v, drop_key, row, col, e.drop);
            if (e.residual) {
              float x[32];
              load32(e.residual, e.res_dtype, (int64_t)row * e.ldres + col, full ? r_vec : 0, ncols, x);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += x[i];
            }
            store32(e.c, e.c_dtype, (int64_t)row * e.ldc + col, full ? c_vec : 0, ncols, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  bind_context();
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

// 2-D bf16 tensor: `inner` contiguous elements, `outer` rows `ld` elements apart; box = 64 x box_rows.
int make_tmap(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, int64_t ld, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("gemm_tcgen05: cuTensorMapEncodeTiled is not available from the driver");
    return FAVIT_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tcgen05: cuTensorMapEncodeTiled failed with CUresult %d (ptr=%p inner=%llu outer=%llu ld=%lld)",
              (int)r, ptr, (unsigned long long)inner, (unsigned long long)outer, (long long)ld);
    return FAVIT_ERR_CUDA;
  }
  return FAVIT_OK;
}

template <int BN>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const KParams& kp, int grid, cudaStream_t st) {
  static bool configured = false;  // per process; the attribute is per function
  const uint32_t smem = smem_bytes_for(BN);
  if (!configured) {
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    configured = true;
  }
  gemm_bf16_tcgen05_kernel<BN><<<grid, kThreads, smem, st>>>(ta, tb, kp);
  FAVIT_CHECK_LAUNCH();
  return FAVIT_OK;
}

}  // namespace

int gemm_bf16(const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, int M, int N, int K,
              const Epilogue& epi, int force_bn, int force_splits, cudaStream_t st) {
  FAVIT_CHECK_ARG(A && B && epi.c, "gemm_tcgen05: null operand");
  FAVIT_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_tcgen05: M,N,K must be positive (got %d,%d,%d)", M, N, K);
  FAVIT_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0), "gemm_tcgen05: operands must be 16-byte aligned");
  FAVIT_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0, "gemm_tcgen05: leading dimensions must be multiples of 8 elements");
  // force_bn: 0 = auto (CTA-pair kernel for large shapes), 512 = CTA-pair kernel, 64/128/256 = this kernel
  if ((force_bn == 0 || force_bn == 512) && gemm_bf16_2cta_applicable(M, N, K, epi, lda, ldb))
    return gemm_bf16_2cta(A, a_mn, lda, B, b_mn, ldb, M, N, K, epi, force_splits, st);
  if (force_bn == 512) {  // the caller asked for the CTA-pair kernel by name: never substitute another one silently
    set_error("gemm_tcgen05: the CTA-pair kernel does not apply to M=%d N=%d K=%d with this epilogue", M, N, K);
    return FAVIT_ERR_UNSUPPORTED;
  }
  const int sms = num_sms();
  const int m_tiles = ceil_div(M, BM);
  const int k_blocks = ceil_div(K, BK);

  // Tile width and split-K factor from a small cost model (cycles per CTA, slowest CTA decides):
  //   a k-block costs 4 MMAs of max(128, BN/2)-ish cycles, but BN < 256 is shared-memory-bandwidth bound
  //   (A is re-read for fewer output columns), so narrower tiles are only worth it when they fill idle SMs.
  const bool can_split = epi.c_dtype == FAVIT_F32 && epi.bias == nullptr && epi.act == FAVIT_EPI_NONE &&
                         epi.residual == nullptr && epi.split_ok;
  int bn = force_bn, splits = force_splits > 0 ? force_splits : 1;
  if (force_bn == 0 || force_splits == 0) {
    const int cand[3] = {256, 128, 64};
    const double kb_cycles[3] = {512.0, 300.0, 200.0};
    const double fixed[3] = {2200.0, 1300.0, 900.0};
    double best = 1e30;
    for (int c = 0; c < 3; ++c) {
      if (force_bn && cand[c] != force_bn) continue;
      if (!force_bn && c > 0 && N <= cand[c]) {
        // a narrower tile than N needs is pointless unless the wider one does not exist
      }
      const int64_t t = (int64_t)m_tiles * ceil_div(N, cand[c]);
      const int smax = force_splits > 0 ? force_splits : (can_split ? min(32, max(1, k_blocks / 4)) : 1);
      for (int sp = (force_splits > 0 ? force_splits : 1); sp <= smax; ++sp) {
        const int kps = ceil_div(k_blocks, sp);
        const int sp_eff = ceil_div(k_blocks, kps);
        const double waves = (double)ceil_div64(t * sp_eff, sms);
        const double cost = waves * (kps * kb_cycles[c] + fixed[c] * (sp_eff > 1 ? 1.5 : 1.0));
        if (cost < best) { best = cost; bn = cand[c]; splits = sp_eff; }
      }
    }
  }
  FAVIT_CHECK_ARG(bn == 64 || bn == 128 || bn == 256, "gemm_tcgen05: BN must be 64, 128 or 256");
  const int n_tiles = ceil_div(N, bn);
  const int64_t tiles = (int64_t)m_tiles * n_tiles;
  if (splits < 1) splits = 1;
  FAVIT_CHECK_ARG(splits == 1 || can_split, "gemm_tcgen05: split-K needs a plain fp32 accumulate epilogue");
  int kb_per_split = ceil_div(k_blocks, splits);
  splits = ceil_div(k_blocks, kb_per_split);

  CUtensorMap ta, tb;
  int rc;
  if (a_mn) rc = make_tmap(&ta, A, (uint64_t)M, (uint64_t)K, lda, 64);
  else rc = make_tmap(&ta, A, (uint64_t)K, (uint64_t)M, lda, BM);
  if (rc) return rc;
  if (b_mn) rc = make_tmap(&tb, B, (uint64_t)N, (uint64_t)K, ldb, 64);
  else rc = make_tmap(&tb, B, (uint64_t)K, (uint64_t)N, ldb, (uint32_t)bn);
  if (rc) return rc;

  KParams kp;
  kp.M = M; kp.N = N; kp.K = K;
  kp.m_tiles = m_tiles; kp.n_tiles = n_tiles; kp.k_blocks = k_blocks;
  kp.splits = splits; kp.kb_per_split = kb_per_split;
  kp.a_mn = a_mn; kp.b_mn = b_mn;
  kp.epi = epi;
  const int64_t units = tiles * splits;
  const int grid = (int)min((int64_t)sms, units);
  note_kernel("gemm_bf16_tcgen05_kernel<BN=%d> act=%d splits=%d a_mn=%d b_mn=%d", bn, epi.act, splits, a_mn, b_mn);
  switch (bn) {
    case 64: return launch<64>(ta, tb, kp, grid, st);
    case 128: return launch<128>(ta, tb, kp, grid, st);
    default: return launch<256>(ta, tb, kp, grid, st);
  }
}

}  // namespace tc
}  // namespace favit
