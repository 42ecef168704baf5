// mm_common.cuh — shared types and device helpers of the search tick (sm_100a); see mm_kernels.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mm_engine.h"

namespace mm {

constexpr int kBlock = 1024;          // threads per CTA for hist / place
constexpr int kJ = 4;                 // batches per round in k_place
constexpr uint32_t kRound = kBlock * kJ;
constexpr uint32_t kNone = 0x1FFFu;   // list terminator (13-bit node ids)
constexpr uint32_t kMaxRows = 2048;   // rows (CTAs) of the histogram matrix
constexpr uint32_t kTile = 2048;      // players per TMA tile in k_place2
constexpr uint32_t kMaxStages = 4;    // depth of the (bin, id) shared-memory ring
constexpr uint32_t kTileBytes = kTile * (8 + 2);
constexpr uint32_t kDenseStride = 66;  // u16 per bin row of the dense group-size matrix (64 batches + pad)
constexpr uint32_t kDenseMaxBins = 256;
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr uint64_t kTombKey = 0xFFFFFFFFFFFFFFFEull;
constexpr uint64_t kFreeVal = 0xFFFFFFFFFFFFFFFFull;
constexpr uint64_t kPending = 0x8000000000000000ull;

struct PoolView {
  uint64_t* id;
  int32_t* rating;
  uint8_t* mode;
  uint8_t* tsize;
  uint32_t* ts;
  uint16_t* bin;  // derived at ingest: mode * stride + lut[clamp(rating)]; K = removed while queued
};

struct BinMap {
  const uint16_t* lut;  // [KR] clamp key -> bin offset inside the mode
  int32_t key_lo;       // rmin - 1
  uint32_t KR;          // rmax - rmin + 3
  uint32_t stride;      // bins per mode
  uint32_t K;           // live bins; bin K = removed-while-queued players
};

struct SegInfo {        // one (mode, group) partition
  uint32_t n;           // alive players
  uint32_t n_lobbies;
  uint32_t member_base; // first slot in member_ids
  uint32_t lobby_base;  // first lobby index
};

struct TickCtr {
  uint32_t gbar;  // grid barrier of the fused tick kernel
  uint32_t n_lobbies, n_matched, n_alive, n_dead, n_resid;
  uint32_t reserved0;
  uint32_t heavy;  // some bin expects > 4 players per tile: use warp-aggregated ranking
  unsigned long long t[8];  // fused kernel: %globaltimer (ns) at phase boundaries, CTA 0; [6],[7]: max over CTAs
};

// Active set slot = {key, value} adjacent in one 16-byte pair: the claim's CAS on the key and atomicMin on the
// value, the winner check and the commit all touch the same 32-byte sector (one DRAM access instead of four).
struct Strided64 {
  unsigned long long* p;
  __device__ __forceinline__ unsigned long long& operator[](uint64_t h) const { return p[2 * h]; }
};
struct ActiveView {
  Strided64 keys;  // keys[h] = kv[2h]
  Strided64 vals;  // vals[h] = kv[2h + 1]
  uint64_t mask;   // capacity - 1, 0 = no active set
};

__device__ __forceinline__ uint64_t hash64(uint64_t x) {
  x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
  return x;
}

// L2 cache-policy hints.  The placement kernel scatters 8-byte ids into member_ids: the
// 4 writes that complete a 32-byte sector arrive at unrelated times, so member_ids has to
// stay L2-resident until the kernel ends (evict_last) while the input columns stream
// through once (evict_first, no L1 allocation).
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ int32_t ld_stream_s32(const int32_t* a, uint64_t pol) {
  int32_t v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u8(const uint8_t* a, uint64_t pol) {
  uint32_t v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* a, uint64_t pol) {
  uint64_t v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_keep_u64(uint64_t* a, uint64_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(a), "l"(v), "l"(pol) : "memory");
}

// ---- TMA (1-D bulk copy) + mbarrier, CTA-local ------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Grid-wide barrier for the fused tick kernel (cooperative launch: all CTAs are co-resident).
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v < target) __nanosleep(32);
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}
// global -> shared bulk copy (SASS: UBLKCP), completion counted on `bar`, L2 evict-first
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

// In-place exclusive scan of a shared-memory array a[0..n) by the whole CTA; returns the
// total.  s_tmp must hold >= 33 words.  Warp-shuffle scan: 3 barriers.
template <int BLOCK>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t* a, uint32_t n, uint32_t* s_tmp) {
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t per = (n + BLOCK - 1) / BLOCK;
  const uint32_t lo = tid * per < n ? tid * per : n, hi = (lo + per < n) ? lo + per : n;
  uint32_t local = 0;
  for (uint32_t i = lo; i < hi; ++i) local += a[i];
  uint32_t incl = local;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= (uint32_t)off) incl += v;
  }
  if (lane == 31) s_tmp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < BLOCK / 32 ? s_tmp[lane] : 0, wi = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, off);
      if (lane >= (uint32_t)off) wi += v;
    }
    s_tmp[lane] = wi - w;                 // exclusive warp offsets
    if (lane == 31) s_tmp[32] = wi;       // grand total
  }
  __syncthreads();
  uint32_t run = s_tmp[warp] + incl - local;
  for (uint32_t i = lo; i < hi; ++i) { const uint32_t v = a[i]; a[i] = run; run += v; }
  const uint32_t total = s_tmp[32];
  __syncthreads();
  return total;
}

__device__ __forceinline__ uint32_t bin_of(const BinMap& bm, const uint16_t* s_lut, int32_t rating, uint32_t mode) {
  if (mode == MM_MODE_DEAD) return bm.K;
  int32_t hi = bm.key_lo + (int32_t)bm.KR - 1;
  int32_t r = rating < bm.key_lo ? bm.key_lo : (rating > hi ? hi : rating);
  return mode * bm.stride + s_lut[r - bm.key_lo];
}

}  // namespace mm
