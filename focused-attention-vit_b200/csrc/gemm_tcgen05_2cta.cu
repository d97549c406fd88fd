// bf16 GEMM, CTA-pair version: tcgen05.mma.cta_group::2 (M = 256 per pair), operands by TMA, epilogue through
// shared-memory staging and TMA stores.  Same contract and reference spans as gemm_tcgen05.cu (mhla.py:100,158 and
// the MLP of vit.py:107-139, forward / dgrad / wgrad); this is the kernel the large shapes of the training step use.
//
// Why a second kernel: with one CTA per tile, a 128 x 256 tile reads 12 KiB of operands from shared memory per 128
// tensor-pipe cycles (96 of the 128 B/clk an SM has) and the per-row global accesses of its epilogue compete for what
// is left.  A CTA pair splits B: each SM stages 128 x 64 of A and only 128 x 64 of B per k-block (64 B/clk), the pair
// issues one M = 256 instruction, and the freed shared-memory bandwidth pays for an epilogue that goes
//   TMEM -> registers (bias / GELU / GELU') -> swizzled shared memory -> cp.async.bulk.tensor store
// so that global traffic is whole 128-byte lines issued by the TMA unit (and `cp.reduce.async.bulk ... add` replaces
// per-element atomics for split-K weight gradients).  The GELU' operand is prefetched by TMA loads as well.
//
// Cluster of 2 CTAs = one 256 x 256 output tile at a time (persistent, 74 clusters).  Cluster c owns the units c,
// c + 74, c + 148, ... (for the training shapes a perfectly balanced schedule, which a greedy global counter is not:
// measured 34.1 against 30.4 ms per ViT-B step).  Two ways to walk that list, chosen per launch
// (favit_set_gemm_tile_scheduler):
//   STEAL = false (default): plain striding, nothing shared.  Fastest when the GPU is ours alone.
//   STEAL = true: the cluster TAKES its units one at a time through an atomic cursor, and a cluster whose own list is
//     used up takes units from the lists of other clusters.  A cluster that becomes resident late - because another
//     kernel (the NCCL all-reduce of data-parallel training) holds some SMs - finds its list already worked off by the
//     others instead of running it as a second wave after everyone else has gone.  Warp 3 of the leader CTA is the
//     scheduler: it draws units ahead of the loads and publishes each to every role of both CTAs through a 4-deep
//     shared-memory ring (st.shared::cluster + mbarriers).  (The producer thread must not do this itself: a
//     cluster-scope release between two tiles delays the next loads by ~0.6 us, which the 5-stage operand pipeline
//     does not absorb - measured 32.5 against 30.8 ms per step.)  The last cluster to finish re-arms the cursors, so
//     launches need no memset.  Costs ~1 % when nothing contends (30.9 against 30.5 ms), hence the switch.
// Per CTA, 384 threads:
//   warp 0: TMA producer (its own 128 rows of A and 128 rows of B; bytes are signalled on the LEADER's full barrier)
//   warp 1: MMA issuer (leader CTA only); tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs
//   warp 2: TMEM allocator (cta_group::2, 512 columns = two 256-column fp32 accumulators)
//   warp 3: tile scheduler (STEAL only; leader CTA only)
//   warps 4-11: epilogue; warp w owns TMEM lanes 32*(w%4).. (32 rows) and 128 of the 256 columns.
#include <atomic>
#include <mutex>

#include "favit_common.cuh"
#include "gemm_epilogue.cuh"
#include "gemm_tcgen05.h"
#include "tcgen05_ptx.cuh"

namespace favit {
namespace tc {
namespace {

using namespace ptx;

std::atomic<int> g_tile_scheduler{0};   // 0 = static striding, 1 = work stealing (favit_set_gemm_tile_scheduler)

constexpr int BMC = 128;   // rows per CTA
constexpr int BM2 = 256;   // rows per cluster tile
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr uint32_t kABytes = BMC * BK * 2;        // 16 KiB
constexpr uint32_t kBBytes = (BN / 2) * BK * 2;   // 16 KiB: each CTA stages half of B
constexpr uint32_t kStage = kABytes + kBBytes;
constexpr uint32_t kSlabBytes = 64 * 128;
constexpr uint32_t kUnit = 32 * 128;              // one staging unit: 32 rows x 128 bytes
constexpr int kSched = 4;                         // depth of the tile-index ring
constexpr int kSchedSlots = 64;                   // sets of global scheduler words, rotated per launch
constexpr int kSchedWords = 128;                  // per set: one list cursor per cluster (<= 96), kClaimed, kDone
constexpr int kClaimed = 96;                      // units taken so far, by anyone (a hint: lets thieves skip the scan)
constexpr int kDone = 97;                         // clusters that have stopped drawing
// readers of a ring slot: both producers, the leader's MMA issuer, the epilogue warps of both CTAs
constexpr uint32_t kSchedReaders = 2 + 1 + 2 * kEpiWarps;

// Zero at module load; the last cluster of every launch puts the words it used back to zero.
__device__ unsigned int g_tile_sched[kSchedSlots][kSchedWords];

template <bool AUX> constexpr int stages() { return AUX ? 4 : 5; }
template <bool AUX> constexpr int out_bufs() { return AUX ? 1 : 2; }
template <bool AUX> constexpr uint32_t smem_bytes() {
  return stages<AUX>() * kStage + kEpiWarps * out_bufs<AUX>() * kUnit + (AUX ? kEpiWarps * 2 * kUnit : 0) + 1024 + 512;
}

struct K2Params {
  int M, N, K;
  int m_tiles, n_tiles, k_blocks, splits, kb_per_split;
  int a_mn, b_mn;
  const float* bias;
  int act;       // FAVIT_EPI_NONE / GELU (writes the pre-activation through tmC2) / DGELU_MUL (reads tmAux)
  int c_fp32;    // C element type: 1 = fp32 (box 32 x 32), 0 = bf16 (box 64 x 32)
  int reduce;    // accumulate into C with TMA reduce-add (split-K or accumulate)
  float* colsum; // [N] fp32, accumulated: column sums of the bf16 output (plain / GELU' epilogues only)
  DropSpec drop; // keep-mask applied after the activation (GELU: to the activation only, not to the saved pre-activation)
  const float* residual;  // fp32 [M,N] added to an fp32 C (the block's fc2 + residual), row pitch ldres; nullptr = none
  int64_t ldres;
  int sched_slot;         // which g_tile_sched set this launch draws its tiles from
};

__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_shared_cluster_u32(uint32_t addr, uint32_t cta, uint32_t v) {
  asm volatile(
      "{\n\t.reg .b32 r;\n\t"
      "mapa.shared::cluster.u32 r, %0, %1;\n\t"
      "st.shared::cluster.u32 [r], %2;\n\t}"
      ::"r"(addr), "r"(cta), "r"(v)
      : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// this lane's 32-column slice -> its row of a swizzled staging unit (bf16: 4 x 16-byte chunks at chunk offset `c4`)
__device__ __forceinline__ void stage_bf16(uint8_t* unit, int lane, int c4, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
    u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
    u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
    u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(unit + lane * 128 + (((c4 + j) ^ (lane & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ void stage_f32(uint8_t* unit, int lane, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(unit + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void unstage_bf16(const uint8_t* unit, int lane, int c4, float (&x)[32]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = *reinterpret_cast<const uint4*>(unit + lane * 128 + (((c4 + j) ^ (lane & 7)) << 4));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      x[8 * j + 2 * t] = __uint_as_float(w[t] << 16);
      x[8 * j + 2 * t + 1] = __uint_as_float(w[t] & 0xffff0000u);
    }
  }
}

template <bool AUX, bool STEAL>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                              const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2,
                              const __grid_constant__ CUtensorMap tmAux, const K2Params p) {
  constexpr int S = stages<AUX>();
  constexpr int NOUT = out_bufs<AUX>();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_u32);           // generic pointer to the aligned base
  const uint32_t out_base = smem_base + S * kStage;
  const uint32_t aux_base = out_base + kEpiWarps * NOUT * kUnit;
  const uint32_t bar_base = aux_base + (AUX ? kEpiWarps * 2 * kUnit : 0);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * S + 2 + a); };
  auto aux_bar = [&](int w) { return bar_base + 8u * (2 * S + 4 + w); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * S + 4 + kEpiWarps);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + (tmem_slot - smem_base));
  auto sfull_bar = [&](int i) { return tmem_slot + 8u * (1 + i); };            // unit index i of the ring is valid
  auto sempty_bar = [&](int i) { return tmem_slot + 8u * (1 + kSched + i); };  // (leader's copy) every reader has it
  auto sched_val = [&](int i) { return tmem_slot + 8u * (1 + 2 * kSched) + 4u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_clusters = gridDim.x >> 1;
  const int total_units = p.m_tiles * p.n_tiles * p.splits;
  unsigned int* const sched = g_tile_sched[p.sched_slot];

  const int cluster_id = blockIdx.x >> 1;
  // units on cluster v's list (v, v + num_clusters, ...); the launch has at most one cluster per unit
  auto list_len = [&](int v) { return (unsigned int)((total_units - v + num_clusters - 1) / num_clusters); };
  // the scheduler thread draws from its own list now: the round trip runs under the barrier / TMEM set-up
  unsigned int drawn = 0;
  if (STEAL && warp == 3 && lane == 0 && leader) drawn = atomicAdd(sched + cluster_id, 1u);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    if (p.act == FAVIT_EPI_GELU) prefetch_tmap(&tmC2);
    if (AUX) prefetch_tmap(&tmAux);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * kEpiWarps);  // every epilogue warp of both CTAs
    }
    for (int w = 0; w < kEpiWarps; ++w) mbar_init(aux_bar(w), 1);
    if (STEAL) {
      for (int i = 0; i < kSched; ++i) {
        mbar_init(sfull_bar(i), 1);
        mbar_init(sempty_bar(i), kSchedReaders);
      }
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // barriers of both CTAs are initialised before anyone signals across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Every role asks for its next unit here; >= total_units when there is none.  Static: the cluster's own list by
  // index.  STEAL: the next entry of the tile ring.
  int static_u = cluster_id - num_clusters;
  int sr_slot = 0;
  uint32_t sr_phase = 0;
  // Hand-back of a slot after `u` was read from it.  Relaxed arrival (nothing is published, see mbar_arrive_relaxed);
  // the branch on the loaded value - never negative - keeps the read of the slot ahead of the arrival that lets the
  // scheduler overwrite it.
  auto took_unit = [&](int u) {
    if (lane == 0 && u >= 0) {
      if (leader) mbar_arrive_relaxed(sempty_bar(sr_slot));
      else mbar_arrive_cluster_relaxed(sempty_bar(sr_slot), 0);
    }
    if (++sr_slot == kSched) { sr_slot = 0; sr_phase ^= 1u; }
  };
  // the slot is written by the leader's scheduler: a remote write for the peer CTA (acquire at cluster scope)
  // `whole_warp`: called by all 32 lanes (epilogue warps) or by lane 0 alone (producer, MMA issuer).
  auto next_unit = [&](bool whole_warp) -> int {
    if (!STEAL) {
      static_u += num_clusters;
      return static_u < total_units ? static_u : total_units;
    }
    if (leader) mbar_wait(sfull_bar(sr_slot), sr_phase);
    else mbar_wait_cluster(sfull_bar(sr_slot), sr_phase);
    const int u = (int)ld_shared_u32(sched_val(sr_slot));
    if (whole_warp) __syncwarp();                      // every lane has read the slot before it is handed back
    took_unit(u);
    return u;
  };
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = next_unit(false); u < total_units; u = next_unit(false)) {
        const int tile = u / p.splits, split = u % p.splits;
        const int m0 = (tile / p.n_tiles) * BM2 + (int)rank * BMC;
        const int n0 = (tile % p.n_tiles) * BN + (int)rank * (BN / 2);
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * kStage;
          const uint32_t sb = sa + kABytes;
          if (leader) mbar_expect_tx(full_bar(stage), 2 * kStage);  // both CTAs' loads land on this barrier
          const int k0 = kb * BK;
          if (p.a_mn) {
            tma_load_2d_pair(sa, &tmA, full_bar(stage), m0, k0);
            tma_load_2d_pair(sa + kSlabBytes, &tmA, full_bar(stage), m0 + 64, k0);
          } else {
            tma_load_2d_pair(sa, &tmA, full_bar(stage), k0, m0);
          }
          if (p.b_mn) {
            tma_load_2d_pair(sb, &tmB, full_bar(stage), n0, k0);
            tma_load_2d_pair(sb + kSlabBytes, &tmB, full_bar(stage), n0 + 64, k0);
          } else {
            tma_load_2d_pair(sb, &tmB, full_bar(stage), k0, n0);
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 3) {
    // ===================== tile scheduler (leader CTA only) =====================
    if (STEAL && leader && lane == 0) {
      int sw_slot = 0;
      uint32_t sw_phase = 0;
      // hand unit `v` to every reader of both CTAs (waits until the slot's previous unit has been read by all of them:
      // that keeps the scheduler about two tiles ahead of the producers, the ring being 4 deep and the epilogue two
      // tiles behind the loads)
      auto publish = [&](int v) {
        mbar_wait_cluster(sempty_bar(sw_slot), sw_phase ^ 1u);
        st_shared_cluster_u32(sched_val(sw_slot), 0, (uint32_t)v);
        st_shared_cluster_u32(sched_val(sw_slot), 1, (uint32_t)v);
        mbar_arrive_cluster(sfull_bar(sw_slot), 0);
        mbar_arrive_cluster(sfull_bar(sw_slot), 1);
        if (++sw_slot == kSched) { sw_slot = 0; sw_phase ^= 1u; }
      };
      const unsigned int own_len = list_len(cluster_id);
      int victim = -1;
      // a unit from another cluster's list, or total_units when every unit of the launch has been taken
      auto steal = [&]() -> int {
        for (;;) {
          if (victim >= 0) {
            const unsigned int t = atomicAdd(sched + victim, 1u);
            if (t < list_len(victim)) {
              atomicAdd(sched + kClaimed, 1u);
              return victim + (int)t * num_clusters;
            }
            victim = -1;
          }
          if (ld_relaxed_gpu(sched + kClaimed) >= (unsigned int)total_units) return total_units;
          // look at the other clusters' cursors, 16 loads in flight at a time, starting behind this cluster
          for (int base = 1; base < num_clusters && victim < 0; base += 16) {
            unsigned int cur[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int v = (cluster_id + base + j) % num_clusters;
              cur[j] = base + j < num_clusters ? ld_relaxed_gpu(sched + v) : 0xffffffffu;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int v = (cluster_id + base + j) % num_clusters;
              if (victim < 0 && base + j < num_clusters && cur[j] < list_len(v)) victim = v;
            }
          }
          if (victim < 0) return total_units;
        }
      };
      // publish() releases at cluster scope, which waits for this thread's outstanding atomics: the draws for the
      // NEXT unit are therefore issued after the publish of this one (a draw in flight across the first publish cost
      // ~2 us at the start of every launch).
      bool own_left = true;
      for (;;) {
        int unit = -1;
        if (own_left) {
          if (drawn < own_len) unit = cluster_id + (int)drawn * num_clusters;
          else own_left = false;
        }
        const bool from_own = unit >= 0;
        if (!from_own) unit = steal();
        publish(unit);
        if (unit >= total_units) break;
        if (from_own) {
          atomicAdd(sched + kClaimed, 1u);
          drawn = atomicAdd(sched + cluster_id, 1u);
        }
      }
      // all draws of this cluster are done; the last cluster to get here re-arms the words for the next launch
      __threadfence();
      if (atomicAdd(sched + kDone, 1u) == (unsigned int)num_clusters - 1u) {
        for (int v = 0; v < num_clusters; ++v) sched[v] = 0u;
        sched[kClaimed] = 0u;
        sched[kDone] = 0u;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc(BM2, BN, p.a_mn, p.b_mn);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = next_unit(false); u < total_units; u = next_unit(false)) {
        const int split = u % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.kb_per_split);
        mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kStage;
          const uint32_t sb = sa + kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = p.a_mn ? make_smem_desc(sa + k * 2048u, kSlabBytes, 1024u)
                                       : make_smem_desc(sa + k * 32u, 16u, 1024u);
            const uint64_t db = p.b_mn ? make_smem_desc(sb + k * 2048u, kSlabBytes, 1024u)
                                       : make_smem_desc(sb + k * 32u, 16u, 1024u);
            umma_bf16_pair(tmem_c, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(empty_bar(stage));
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        umma_commit_pair(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs) =====================
    const int ew = warp - 4;
    const int wq = warp & 3;      // TMEM lane quarter
    const int chalf = ew >> 2;    // which 128 of the tile's 256 columns
    uint8_t* outbuf = smem_gen + (out_base - smem_base) + ew * NOUT * kUnit;
    const uint32_t outbuf_u32 = out_base + ew * NOUT * kUnit;
    uint8_t* auxbuf = smem_gen + (aux_base - smem_base) + ew * 2 * kUnit;
    const uint32_t auxbuf_u32 = aux_base + ew * 2 * kUnit;
    const bool b_vec = p.bias ? (((uintptr_t)p.bias) % 16 == 0) : false;
    const unsigned long long drop_key = p.drop.seed ? __ldg(p.drop.seed) + p.drop.offset : 0ull;
    int acc = 0;
    uint32_t acc_phase = 0, aux_phase = 0;
    int obuf = 0;  // next staging buffer (round robin)
    for (int u = next_unit(true); u < total_units; u = next_unit(true)) {
      const int tile = u / p.splits;
      const int row0 = (tile / p.n_tiles) * BM2 + (int)rank * BMC + wq * 32;
      const int col0 = (tile % p.n_tiles) * BN + chalf * 128;
      if (AUX && lane == 0) {  // GELU' operand of this warp's 32 x 128 patch, prefetched while the MMAs still run
        mbar_expect_tx(aux_bar(ew), 2 * kUnit);
        tma_load_2d(auxbuf_u32, &tmAux, aux_bar(ew), col0, row0);
        tma_load_2d(auxbuf_u32 + kUnit, &tmAux, aux_bar(ew), col0 + 64, row0);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (AUX) {
        mbar_wait(aux_bar(ew), aux_phase);
        aux_phase ^= 1u;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(acc * BN + chalf * 128);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_wait_ld();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        const int col = col0 + c * 32;
        if (p.bias) {
          if (b_vec && col + 32 <= p.N) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + i);
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col + i < p.N) v[i] += __ldg(p.bias + col + i);
          }
        }
        if (p.drop.seed && p.act != FAVIT_EPI_GELU && !AUX && !p.reduce) dropout32(v, drop_key, row0 + lane, col, p.drop);
        if (p.residual && p.c_fp32 && !p.reduce && row0 + lane < p.M && col + 32 <= p.N) {
          // fp32 residual stream: this lane's 32 columns are one 128-byte line of its row
          const float4* rp = reinterpret_cast<const float4*>(p.residual + (int64_t)(row0 + lane) * p.ldres + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 r4 = __ldg(rp + i);
            v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w;
          }
        }
        if (p.c_fp32) {
          // one staging unit per 32 columns
          uint8_t* ub = outbuf + obuf * kUnit;
          if (lane == 0) tma_store_wait_read<NOUT - 1>();
          __syncwarp();
          stage_f32(ub, lane, v);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && col < p.N) {
            if (p.reduce) tma_reduce_add_2d(&tmC, outbuf_u32 + obuf * kUnit, col, row0);
            else tma_store_2d(&tmC, outbuf_u32 + obuf * kUnit, col, row0);
          }
          if (lane == 0) tma_store_commit();
          obuf = (obuf + 1) % NOUT;
        } else if (p.act == FAVIT_EPI_GELU) {
          // two bf16 outputs per 64-column unit: buffer 0 = pre-activation, buffer 1 = activation.  Each has its own
          // bulk group, so the pre-activation store of this unit runs under its GELU arithmetic and the wait for the
          // previous unit's activation store comes after the first half's arithmetic instead of before any of it.
          if ((c & 1) == 0) {
            if (lane == 0) tma_store_wait_read<1>();  // the previous unit's pre-activation store has read buffer 0
            __syncwarp();
          }
          stage_bf16(outbuf, lane, (c & 1) * 4, v);
          if (c & 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (col - 32 < p.N) tma_store_2d(&tmC2, outbuf_u32, col - 32, row0);
              tma_store_commit();
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
          if (p.drop.seed) dropout32(v, drop_key, row0 + lane, col, p.drop);
          if ((c & 1) == 0) {
            if (lane == 0) tma_store_wait_read<0>();  // ... and its activation store has read buffer 1
            __syncwarp();
          }
          stage_bf16(outbuf + (NOUT - 1) * kUnit, lane, (c & 1) * 4, v);
          if (c & 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (col - 32 < p.N) tma_store_2d(&tmC, outbuf_u32 + (NOUT - 1) * kUnit, col - 32, row0);
              tma_store_commit();
            }
          }
        } else {
          if (AUX) {  // v *= gelu'(pre-activation)
            float x[32];
            unstage_bf16(auxbuf + (c >> 1) * kUnit, lane, (c & 1) * 4, x);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= dgelu_fast(x[i]);
            if (p.drop.seed) dropout32(v, drop_key, row0 + lane, col, p.drop);
          }
          uint8_t* ub = outbuf + obuf * kUnit;
          if ((c & 1) == 0) {
            if (lane == 0) tma_store_wait_read<NOUT - 1>();
            __syncwarp();
          }
          stage_bf16(ub, lane, (c & 1) * 4, v);
          if (c & 1) {
            fence_proxy_async_smem();
            __syncwarp();
            const int ucol = col - 32;
            if (lane == 0 && ucol < p.N) tma_store_2d(&tmC, outbuf_u32 + obuf * kUnit, ucol, row0);
            if (lane == 0) tma_store_commit();
            if (p.colsum) {
              // column sums of the staged 32 x 64 bf16 unit (the values exactly as stored): the bias gradient of the
              // layer that consumes this gradient, without a second pass over the tensor.  Lane l owns columns 2l, 2l+1;
              // rows past M hold exact zeros (TMA zero-fills A), so they add nothing.
              float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) {
                const uint32_t w2 = *reinterpret_cast<const uint32_t*>(ub + r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) +
                                                                       (lane & 3) * 4);
                s0 += __uint_as_float(w2 << 16);
                s1 += __uint_as_float(w2 & 0xffff0000u);
              }
              const int cc = ucol + 2 * lane;
              if (cc < p.N) atomicAdd(p.colsum + cc, s0);
              if (cc + 1 < p.N) atomicAdd(p.colsum + cc + 1, s1);
            }
            obuf = (obuf + 1) % NOUT;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive_relaxed(tempty_bar(acc));
        else mbar_arrive_cluster_relaxed(tempty_bar(acc), 0);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) tma_store_wait_all<0>();  // the stores read this CTA's shared memory: finish before exit
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync();  // the peer may still be signalling our barriers / reading our operands until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// 2-D tensor, `inner` contiguous elements of `esize` bytes, rows `ld` elements apart; box = box_inner x box_rows
int make_tmap(CUtensorMap* tm, const void* ptr, int esize, uint64_t inner, uint64_t outer, int64_t ld,
              uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("gemm_tcgen05_2cta: cuTensorMapEncodeTiled is not available from the driver");
    return FAVIT_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tcgen05_2cta: cuTensorMapEncodeTiled failed with CUresult %d (ptr=%p inner=%llu outer=%llu ld=%lld)",
              (int)r, ptr, (unsigned long long)inner, (unsigned long long)outer, (long long)ld);
    return FAVIT_ERR_CUDA;
  }
  return FAVIT_OK;
}

template <bool AUX, bool STEAL>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc_, const CUtensorMap& tc2,
           const CUtensorMap& taux, const K2Params& kp, int clusters, cudaStream_t st) {
  static bool configured = false;
  constexpr uint32_t smem = smem_bytes<AUX>();
  if (!configured) {
    FAVIT_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_2cta_kernel<AUX, STEAL>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FAVIT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_2cta_kernel<AUX, STEAL>, ta, tb, tc_, tc2, taux, kp));
  count_launch();
  return FAVIT_OK;
}

}  // namespace

bool gemm_bf16_2cta_applicable(int M, int N, int K, const Epilogue& e, int64_t lda, int64_t ldb) {
  // small problems keep the 1-CTA kernel's narrower tiles; "large" = at least one full wave of 256 x 256 x 512 work
  if (M < 256 || N < 256 || K < 64) return false;
  if ((int64_t)ceil_div(M, 256) * ceil_div(N, 256) * ceil_div(K, 64) < 74 * 8) return false;
  // residual: fp32 into an fp32 C, whole 32-column groups (N % 32 == 0) of 16-byte aligned rows; else the 1-CTA kernel
  if (e.residual != nullptr && (e.res_dtype != FAVIT_F32 || e.c_dtype != FAVIT_F32 || N % 32 != 0 || e.ldres % 4 != 0 ||
                                ((uintptr_t)e.residual % 16) != 0 || e.accumulate))
    return false;
  const int es = e.c_dtype == FAVIT_BF16 ? 2 : 4;
  if (((uintptr_t)e.c % 16) || ((e.ldc * es) % 16)) return false;   // TMA store pitch
  if (e.act == FAVIT_EPI_GELU && (e.c_dtype != FAVIT_BF16 || !e.aux_out || ((uintptr_t)e.aux_out % 16) || ((e.ldaux * 2) % 16)))
    return false;
  if (e.act == FAVIT_EPI_DGELU_MUL && (e.c_dtype != FAVIT_BF16 || !e.aux || ((uintptr_t)e.aux % 16) || ((e.ldaux * 2) % 16)))
    return false;
  if (e.accumulate && e.c_dtype != FAVIT_F32) return false;
  if (e.colsum && (e.c_dtype != FAVIT_BF16 || e.act == FAVIT_EPI_GELU)) return false;
  return true;
}

int gemm_bf16_2cta(const void* A, int a_mn, int64_t lda, const void* B, int b_mn, int64_t ldb, int M, int N, int K,
                   const Epilogue& epi, int force_splits, cudaStream_t st) {
  const int sms = num_sms();
  const int clusters_max = sms / 2;
  const int m_tiles = ceil_div(M, BM2), n_tiles = ceil_div(N, BN), k_blocks = ceil_div(K, BK);
  const int64_t tiles = (int64_t)m_tiles * n_tiles;
  const bool can_split = epi.c_dtype == FAVIT_F32 && epi.bias == nullptr && epi.act == FAVIT_EPI_NONE && epi.split_ok &&
                         epi.residual == nullptr;
  int splits = 1;
  if (force_splits > 0) {
    splits = force_splits;
  } else if (can_split) {
    double best = 1e30;
    const int smax = min(32, max(1, k_blocks / 4));
    for (int sp = 1; sp <= smax; ++sp) {
      const int kps = ceil_div(k_blocks, sp);
      const int sp_eff = ceil_div(k_blocks, kps);
      const double waves = (double)ceil_div64(tiles * sp_eff, clusters_max);
      const double cost = waves * (kps * 512.0 + 2500.0);
      if (cost < best) { best = cost; splits = sp_eff; }
    }
  }
  FAVIT_CHECK_ARG(splits == 1 || can_split, "gemm_tcgen05_2cta: split-K needs a plain fp32 accumulate epilogue");
  const int kb_per_split = ceil_div(k_blocks, splits);
  splits = ceil_div(k_blocks, kb_per_split);

  CUtensorMap ta, tb, tcm, tc2, taux;
  int rc;
  rc = a_mn ? make_tmap(&ta, A, 2, (uint64_t)M, (uint64_t)K, lda, 64, 64) : make_tmap(&ta, A, 2, (uint64_t)K, (uint64_t)M, lda, 64, BMC);
  if (rc) return rc;
  rc = b_mn ? make_tmap(&tb, B, 2, (uint64_t)N, (uint64_t)K, ldb, 64, 64) : make_tmap(&tb, B, 2, (uint64_t)K, (uint64_t)N, ldb, 64, BN / 2);
  if (rc) return rc;
  const bool c_fp32 = epi.c_dtype == FAVIT_F32;
  rc = make_tmap(&tcm, epi.c, c_fp32 ? 4 : 2, (uint64_t)N, (uint64_t)M, epi.ldc, c_fp32 ? 32 : 64, 32);
  if (rc) return rc;
  tc2 = tcm;
  taux = tcm;
  if (epi.act == FAVIT_EPI_GELU) {
    rc = make_tmap(&tc2, epi.aux_out, 2, (uint64_t)N, (uint64_t)M, epi.ldaux, 64, 32);
    if (rc) return rc;
  }
  const bool aux = epi.act == FAVIT_EPI_DGELU_MUL;
  if (aux) {
    rc = make_tmap(&taux, epi.aux, 2, (uint64_t)N, (uint64_t)M, epi.ldaux, 64, 32);
    if (rc) return rc;
  }
  K2Params kp;
  kp.M = M; kp.N = N; kp.K = K;
  kp.m_tiles = m_tiles; kp.n_tiles = n_tiles; kp.k_blocks = k_blocks;
  kp.splits = splits; kp.kb_per_split = kb_per_split;
  kp.a_mn = a_mn; kp.b_mn = b_mn;
  kp.bias = epi.bias;
  kp.act = epi.act;
  kp.c_fp32 = c_fp32 ? 1 : 0;
  kp.reduce = (splits > 1 || epi.accumulate) ? 1 : 0;
  kp.colsum = epi.colsum;
  kp.drop = epi.drop;
  kp.residual = (const float*)epi.residual;
  kp.ldres = epi.ldres;
  // Launches that could overlap (different streams, parallel graph branches) must not share a set of scheduler words; a
  // launch re-arms its set when it ends, so stream-ordered launches could share one.  64 sets, round robin: concurrent
  // launches must be fewer than 64 launches apart, and a captured launch keeps its set (include/favit.h states both).
  static std::atomic<unsigned int> next_sched_slot{0};
  const bool steal = g_tile_scheduler.load(std::memory_order_relaxed) == 1;
  kp.sched_slot = steal ? (int)(next_sched_slot.fetch_add(1u, std::memory_order_relaxed) % kSchedSlots) : 0;
  const int clusters = (int)min((int64_t)min(clusters_max, kClaimed), tiles * splits);
  note_kernel("gemm_bf16_tcgen05_2cta_kernel<AUX=%d> act=%d c_fp32=%d reduce=%d splits=%d colsum=%d a_mn=%d b_mn=%d residual=%d steal=%d",
              aux ? 1 : 0, kp.act, kp.c_fp32, kp.reduce, splits, kp.colsum ? 1 : 0, a_mn, b_mn, kp.residual ? 1 : 0,
              steal ? 1 : 0);
  if (steal)
    return aux ? launch<true, true>(ta, tb, tcm, tc2, taux, kp, clusters, st)
               : launch<false, true>(ta, tb, tcm, tc2, taux, kp, clusters, st);
  return aux ? launch<true, false>(ta, tb, tcm, tc2, taux, kp, clusters, st)
             : launch<false, false>(ta, tb, tcm, tc2, taux, kp, clusters, st);
}

int gemm_tile_scheduler(int mode) {
  if (mode == 0 || mode == 1) g_tile_scheduler.store(mode, std::memory_order_relaxed);
  return g_tile_scheduler.load(std::memory_order_relaxed);
}

}  // namespace tc
}  // namespace favit
