// mm_kernels.cuh — device code of the search tick (sm_100a).
//
// The tick replaces, for every queued player at once, the per-request loop of
// Search.Worker.consume/5 (reference matchmaking/lib/search/worker.ex:291-324).
// Under the serialized oracle (oracle/mm_oracle.c) that loop has the closed form
//   "drop inactive players, stable-partition the feed order by (mode, group), cut
//    each partition into lobbies of L"
// which on the GPU is ONE stable counting sort over a small key domain:
//   bin(player) = mode * stride + lut[clamp(rating)]          (K bins, K ~ 5k * modes)
// followed by a per-(mode, group)-partition cut.  One cooperative launch, k_tick<512>, runs the four phases
// (each also exists as a stand-alone kernel):
//   k_hist3    row histograms M[row][bin] from the resident 16-bit bin column (2 B/player, TMA ring)
//   k_colscan  column prefix of M + the tail: per bin, how many players are matched (a prefix of the bin) and the
//              member slot of the first one — policy S0 (reference behaviour) or S1 (rating window, extension)
//   k_place2   stable rank inside the row -> final lobby-major slot; scatters player_id straight to
//              member_ids (reads 10 B/player, writes 8 B); players past their bin's prefix: one bit in left_bits
//   k_epilogue leftover players -> compacted pool (enqueue order kept, work split by rank) + lobby headers
// k_hist / k_place<0|1> are the round's first versions, kept as on-device cross-checks (rank_impl 0/1).
// Integer/HBM-bound work: no tensor cores (BASELINE.json north_star).
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

// ---------------------------------------------------------------------------------------
// k_hist: M[row][bin] = number of the row's players in that bin, and the 16-bit bin column
// bins16[] that k_place2 streams instead of re-deriving bins from rating + mode.
// Coalesced 128-bit rating loads (4 players per thread, 4 such loads in flight), 32-bit
// mode loads, 64-bit bin stores.
// ---------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_hist(PoolView p, uint32_t n, uint32_t chunk, BinMap bm, uint32_t Kp,
                                                uint32_t* __restrict__ M, uint32_t* __restrict__ tot,
                                                uint16_t* __restrict__ bins16) {
  constexpr uint32_t kBlock = BLOCK;
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* hist = smem;
  uint16_t* s_lut = reinterpret_cast<uint16_t*>(hist + Kp);
  const uint32_t tid = threadIdx.x;
  for (uint32_t i = tid; i < Kp; i += kBlock) hist[i] = 0;
  for (uint32_t i = tid; i < bm.KR; i += kBlock) s_lut[i] = bm.lut[i];
  __syncthreads();
  const uint64_t beg64 = (uint64_t)blockIdx.x * chunk;
  const uint32_t beg = beg64 < n ? (uint32_t)beg64 : n;
  const uint32_t end = (beg64 + chunk < n) ? (uint32_t)(beg64 + chunk) : n;
  constexpr int U = 4;
  for (uint32_t i0 = beg + tid * 4; i0 < end; i0 += kBlock * 4 * U) {
    int4 r[U];
    uint32_t m[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i = i0 + u * kBlock * 4;
      if (i + 4 <= end) {
        r[u] = __ldcs(reinterpret_cast<const int4*>(p.rating + i));
        m[u] = __ldcs(reinterpret_cast<const uint32_t*>(p.mode + i));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i = i0 + u * kBlock * 4;
      if (i + 4 <= end) {
        const uint32_t b0 = bin_of(bm, s_lut, r[u].x, m[u] & 0xFF), b1 = bin_of(bm, s_lut, r[u].y, (m[u] >> 8) & 0xFF);
        const uint32_t b2 = bin_of(bm, s_lut, r[u].z, (m[u] >> 16) & 0xFF), b3 = bin_of(bm, s_lut, r[u].w, m[u] >> 24);
        atomicAdd(&hist[b0], 1u); atomicAdd(&hist[b1], 1u); atomicAdd(&hist[b2], 1u); atomicAdd(&hist[b3], 1u);
        if (bins16) *reinterpret_cast<uint2*>(bins16 + i) = make_uint2(b0 | (b1 << 16), b2 | (b3 << 16));
      } else if (i < end) {
        for (uint32_t e = i; e < end; ++e) {
          const uint32_t bb = bin_of(bm, s_lut, p.rating[e], p.mode[e]);
          atomicAdd(&hist[bb], 1u);
          if (bins16) bins16[e] = (uint16_t)bb;
        }
      }
    }
  }
  __syncthreads();
  uint32_t* row = M + (size_t)blockIdx.x * Kp;
  for (uint32_t i = tid; i < Kp; i += kBlock) {
    const uint32_t v = hist[i];
    row[i] = v;
    if (v) atomicAdd(&tot[i], v);  // bin totals (tot[] is zeroed by the previous tick's epilogue)
  }
}

// ---------------------------------------------------------------------------------------
// hist3_body<BLOCK>: row histogram straight from the resident 16-bit bin column (maintained
// at ingest by k_enq_append / k_remove / the epilogue's compaction), streamed through a TMA
// ring of 4 096-player (8 KB) tiles: the tick never touches rating / mode.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kBTile = 4096;
constexpr uint32_t kBTileBytes = kBTile * 2;

template <int BLOCK>
__device__ __forceinline__ void hist3_body(unsigned char* smem_raw, const uint16_t* __restrict__ bins16, uint32_t n,
                                           uint32_t chunk, uint32_t Kp, uint32_t stages, uint32_t* __restrict__ M,
                                           uint32_t* __restrict__ tot) {
  uint16_t* ring = reinterpret_cast<uint16_t*>(smem_raw);                                      // [stages][kBTile]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * kBTileBytes);       // [kMaxStages]
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + (size_t)stages * kBTileBytes + 64);  // [Kp]
  const uint32_t tid = threadIdx.x;
  const uint64_t pol_in = policy_evict_first();
  const uint64_t beg64 = (uint64_t)blockIdx.x * chunk;
  const uint32_t beg = beg64 < n ? (uint32_t)beg64 : n;
  const uint32_t end = (beg64 + chunk < n) ? (uint32_t)(beg64 + chunk) : n;
  const uint32_t n_tiles = (end - beg + kBTile - 1) / kBTile;
  if (tid == 0) {
    for (uint32_t s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0)
    for (uint32_t t = 0; t < stages && t < n_tiles; ++t) {
      mbar_expect_tx(&full[t], kBTileBytes);
      tma_load_1d(ring + (size_t)t * kBTile, bins16 + beg + (size_t)t * kBTile, kBTileBytes, &full[t], pol_in);
    }
  for (uint32_t i = tid; i < Kp; i += BLOCK) hist[i] = 0;
  __syncthreads();
  uint32_t st = 0, parity = 0;
  for (uint32_t t = 0; t < n_tiles; ++t) {
    const uint32_t valid = end - (beg + t * kBTile);
    const uint16_t* tb = ring + (size_t)st * kBTile;
    mbar_wait(&full[st], parity);
#pragma unroll
    for (uint32_t q = tid; q < kBTile / 8; q += BLOCK) {  // 8 bins (128 bits) per thread per step
      const uint32_t o = q * 8;
      if (o + 8 <= valid) {
        const uint4 v = *reinterpret_cast<const uint4*>(tb + o);
        atomicAdd(&hist[v.x & 0xFFFFu], 1u); atomicAdd(&hist[v.x >> 16], 1u);
        atomicAdd(&hist[v.y & 0xFFFFu], 1u); atomicAdd(&hist[v.y >> 16], 1u);
        atomicAdd(&hist[v.z & 0xFFFFu], 1u); atomicAdd(&hist[v.z >> 16], 1u);
        atomicAdd(&hist[v.w & 0xFFFFu], 1u); atomicAdd(&hist[v.w >> 16], 1u);
      } else {
        for (uint32_t k = o; k < valid; ++k) atomicAdd(&hist[tb[k]], 1u);
      }
    }
    __syncthreads();
    if (tid == 0 && t + stages < n_tiles) {
      mbar_expect_tx(&full[st], kBTileBytes);
      tma_load_1d(ring + (size_t)st * kBTile, bins16 + beg + (size_t)(t + stages) * kBTile, kBTileBytes, &full[st], pol_in);
    }
    if (++st == stages) { st = 0; parity ^= 1u; }
  }
  uint32_t* row = M + (size_t)blockIdx.x * Kp;
  for (uint32_t i = tid; i < Kp; i += BLOCK) {
    const uint32_t v = hist[i];
    row[i] = v;
    if (v) atomicAdd(&tot[i], v);
  }
  if (tid == 0)
    for (uint32_t s = 0; s < stages; ++s) mbar_inval(&full[s]);
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK == 512 ? 2 : 1))
    k_hist3(const uint16_t* __restrict__ bins16, uint32_t n, uint32_t chunk, uint32_t Kp, uint32_t stages,
            uint32_t* __restrict__ M, uint32_t* __restrict__ tot) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  hist3_body<BLOCK>(smem_raw, bins16, n, chunk, Kp, stages, M, tot);
}

// ---------------------------------------------------------------------------------------
// k_colscan: exclusive prefix down every column of M.  A column CTA is 32 bins wide
// (lanes = consecutive bins, coalesced) and 16 row-slices deep (warps): every thread sums
// its slice of rows, the slices are scanned through shared memory, then the slice is
// rewritten as running prefixes — one round trip of latency instead of R.
// The LAST CTA of the grid runs concurrently as the "tail": bin totals (accumulated by
// k_hist with global reductions) -> sorted position of every bin -> how many players of every
// bin are matched under the tick's policy (always a PREFIX of the bin in enqueue order) ->
//   outbase[v] = member slot of bin v's first player (exclusive scan of the matched counts)
//   binlim[v]  = outbase[v] + matched players of bin v; a player at or past it stays queued.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kMaxSegs = MM_MAX_GROUPS * MM_MAX_MODES;
constexpr int kScanBlock = 512;
constexpr uint32_t kColScratchWords = (kScanBlock / 32) * 33;      // column CTA scratch
constexpr uint32_t kTailScratchWords = 64 + 4 + 4 * kMaxSegs + 4;  // tail CTA scratch, fixed part

// one 32-bin column group: exclusive prefix down the rows of M (scratch: kColScratchWords)
__device__ __forceinline__ void colscan_cols_body(uint32_t* scratch, uint32_t group, uint32_t R, uint32_t Kp,
                                                  uint32_t* __restrict__ M) {
  uint32_t(*s_part)[33] = reinterpret_cast<uint32_t(*)[33]>(scratch);
  const uint32_t tid = threadIdx.x, x = tid & 31, y = tid >> 5;
  constexpr uint32_t NY = kScanBlock / 32;
  const uint32_t b = group * 32 + x;
  const uint32_t rp = (R + NY - 1) / NY;
  const uint32_t r0 = y * rp < R ? y * rp : R, r1 = (r0 + rp < R) ? r0 + rp : R;
  constexpr int kU = 8;  // independent loads in flight per thread
  uint32_t sum = 0;
  if (b < Kp)
    for (uint32_t r = r0; r < r1; r += kU) {
      uint32_t v[kU];
#pragma unroll
      for (int k = 0; k < kU; ++k) v[k] = (r + k < r1) ? __ldcg(M + (size_t)(r + k) * Kp + b) : 0u;
#pragma unroll
      for (int k = 0; k < kU; ++k) sum += v[k];
    }
  s_part[y][x] = sum;
  __syncthreads();
  uint32_t run = 0;
  for (uint32_t yy = 0; yy < y; ++yy) run += s_part[yy][x];
  if (b < Kp)
    for (uint32_t r = r0; r < r1; r += kU) {
      uint32_t v[kU];
#pragma unroll
      for (int k = 0; k < kU; ++k) v[k] = (r + k < r1) ? __ldcg(M + (size_t)(r + k) * Kp + b) : 0u;
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        if (r + k < r1) M[(size_t)(r + k) * Kp + b] = run;
        run += v[k];
      }
    }
  __syncthreads();  // scratch may be reused by the next group
}

// arguments of the tail (shared by k_colscan and the fused k_tick)
struct TailArgs {
  uint32_t Kp, K, n_segs;
  uint32_t layout;                    // bit 0: matched counts in shared memory; bit 1: bin keys too (tail_words)
  int32_t max_spread;                 // < 0: unlimited (policy S0); >= 0: policy S1, rating order only
  const uint32_t* tot;                // [Kp] bin totals (k_hist)
  const uint32_t* seg_bin_lo;         // [n_segs + 1]
  const uint32_t* seg_L;              // [n_segs]
  const uint16_t* bin_seg;            // [Kp] bin -> segment
  const uint16_t* bin_key;            // [Kp] bin -> clamp key (rating order: ascending inside a segment)
  uint32_t* outbase;                  // [Kp] out: member slot of the bin's first player
  uint32_t* binlim;                   // [Kp] out: outbase + matched players of the bin
  SegInfo* seg;                       // [n_segs] out
  TickCtr* ctr;
};
// shared-memory words of the tail for a layout: bases | matched counts (bit 0) | keys (bit 1)
__host__ __device__ constexpr uint32_t tail_words(uint32_t Kp, uint32_t layout) {
  return kTailScratchWords + (Kp + 2) + ((layout & 1u) ? (Kp + 2) : 0u) + ((layout & 2u) ? (Kp + 3) / 2 : 0u);
}

// The tail.  One pipeline for both policies:
//   bin totals -> bases s_bb -> matched prefix of every bin -> member slot of the bin's first player.
// S0 (reference behaviour): a (mode, group) partition of n players emits floor(n/L) lobbies, the n mod L
//   highest-ranked players stay queued: member slot = sorted position - leftovers of earlier partitions,
//   clipped at the partition's matched end (closed form, only a scan over the partitions).
// S1 (extension): greedy windowed walk over the partition (oracle: orc_run_windowed).  Players of one bin have
//   the same key, so the walk runs on the histogram: from position cur in bin v, lobbies are seeded at
//   cur, cur+L, ... while the seed is still in bin v and its L-th player has key <= key_v + W; whatever is
//   left of bin v afterwards cannot seed and stays queued.  Two-pointer over the bins of the segment.
// Very large key domains (layout bit 0 clear) park m_v in global memory and scan it in place of the bases.
__device__ __forceinline__ void colscan_tail_body(uint32_t* scratch, const TailArgs t) {
  constexpr uint32_t NW = kScanBlock / 32;
  uint32_t* s_tmp = scratch;            // [64]
  uint32_t* s_misc = scratch + 64;      // [4] fullest bin
  uint32_t* s_a = scratch + 68;         // [kMaxSegs] leftovers of earlier segments / member base
  uint32_t* s_lob = s_a + kMaxSegs;     // [kMaxSegs] lobbies of earlier segments
  uint32_t* s_nl = s_lob + kMaxSegs;    // [kMaxSegs] lobbies of the segment
  uint32_t* s_lo = s_nl + kMaxSegs;     // [kMaxSegs + 1] first bin of the segment
  uint32_t* s_bb = scratch + kTailScratchWords;  // [Kp + 1] sorted position of the bin's first player
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, Kp = t.Kp, K = t.K, n_segs = t.n_segs;
  if (tid == 0) s_misc[0] = 0;
  for (uint32_t sg = tid; sg <= n_segs; sg += kScanBlock) s_lo[sg] = t.seg_bin_lo[sg];
  __syncthreads();
  uint32_t lmax = 0;
  for (uint32_t i = tid; i < Kp; i += kScanBlock) {  // coalesced, independent loads
    const uint32_t v = __ldcg(&t.tot[i]);
    s_bb[i] = v;
    if (i < K && v > lmax) lmax = v;
  }
  lmax = __reduce_max_sync(0xFFFFFFFFu, lmax);
  if ((tid & 31) == 0 && lmax) atomicMax(&s_misc[0], lmax);
  __syncthreads();
  const uint32_t total = block_excl_scan<kScanBlock>(s_bb, Kp, s_tmp);
  if (tid == 0) s_bb[Kp] = total;
  __syncthreads();
  const uint32_t alive = s_bb[K], dead = total - alive;
  uint32_t n_matched, tot_lob;

  if (t.max_spread < 0) {
    // S0: lobbies_s = n_s / L; the partition's first lobbies_s * L sorted positions are matched.
    for (uint32_t sg = tid; sg < n_segs; sg += kScanBlock) {
      const uint32_t ns = s_bb[s_lo[sg + 1]] - s_bb[s_lo[sg]], nl = ns / t.seg_L[sg];
      s_nl[sg] = nl; s_lob[sg] = nl; s_a[sg] = ns - nl * t.seg_L[sg];
      t.seg[sg].n = ns; t.seg[sg].n_lobbies = nl;
    }
    __syncthreads();
    const uint32_t n_left = block_excl_scan<kScanBlock>(s_a, n_segs, s_tmp);   // -> leftovers of earlier segments
    tot_lob = block_excl_scan<kScanBlock>(s_lob, n_segs, s_tmp);               // -> lobbies of earlier segments
    n_matched = alive - n_left;
    for (uint32_t sg = warp; sg < n_segs; sg += NW) {  // one warp per partition: no bin -> segment lookups
      const uint32_t lo = s_lo[sg], hi = s_lo[sg + 1], start = s_bb[lo], shift = s_a[sg];
      const uint32_t mend = start + s_nl[sg] * t.seg_L[sg];  // end of the partition's matched positions
      for (uint32_t v = lo + lane; v < hi; v += 32) {
        const uint32_t b0 = s_bb[v], b1 = s_bb[v + 1];
        t.outbase[v] = (b0 < mend ? b0 : mend) - shift;
        t.binlim[v] = (b1 < mend ? b1 : mend) - shift;
      }
      if (lane == 0) { t.seg[sg].member_base = start - shift; t.seg[sg].lobby_base = s_lob[sg]; }
    }
    for (uint32_t v = K + tid; v < Kp; v += kScanBlock) { t.outbase[v] = n_matched; t.binlim[v] = n_matched; }
  } else {
    // S1: greedy windowed walk on the histogram, one thread per partition (see above), then a second scan.
    const bool m_smem = (t.layout & 1u) != 0, key_smem = (t.layout & 2u) != 0;
    uint32_t* s_m = m_smem ? s_bb + Kp + 2 : t.binlim;  // [Kp + 1] matched players of the bin
    uint16_t* s_key = reinterpret_cast<uint16_t*>(s_bb + (m_smem ? 2 : 1) * (Kp + 2));
    const uint16_t* keys = key_smem ? s_key : t.bin_key;
    if (key_smem)
      for (uint32_t v = tid; v < K; v += kScanBlock) s_key[v] = t.bin_key[v];
    for (uint32_t v = K + tid; v < Kp; v += kScanBlock) s_m[v] = 0;
    __syncthreads();
    const int32_t W = t.max_spread;
    // One warp walks TWO partitions at a time (two independent carry chains in flight).  Per bin, off the chain:
    //   reach  = sorted position where keys exceed key_v + W (binary search over the partition's keys)
    //   rsel   = min(reach, b1 - 1 + L): the seeds of bin v are cur, cur + L, ... < min(b1, reach - L + 1), so with
    //            a = rsel - cur the bin seeds a / L lobbies and p2 = rsel - a mod L is the next unconsumed position
    // and on the chain only: cur = max(pos, b0); a; a mod L by a reciprocal multiply; p2; select.  Empty bins
    // fall out of the same arithmetic (cur >= b1), so the 32 bins of a batch are visited by an unrolled loop.
    for (uint32_t sg0 = warp; sg0 < n_segs; sg0 += 2 * NW) {
      uint32_t lo[2], hi[2], L[2], Mrec[2], pos[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const uint32_t sg = sg0 + q * NW;
        const bool on = sg < n_segs;
        lo[q] = on ? s_lo[sg] : 0u; hi[q] = on ? s_lo[sg + 1] : 0u; L[q] = on ? t.seg_L[sg] : 1u;
        Mrec[q] = 0xFFFFFFFFu / L[q];  // umulhi(a, Mrec) is a / L or a / L - 1 for every 32-bit a
        pos[q] = s_bb[lo[q]];
        if (on && lane == 0) t.seg[sg].n = s_bb[hi[q]] - s_bb[lo[q]];
      }
      const uint32_t span0 = hi[0] - lo[0], span1 = hi[1] - lo[1], span = span0 > span1 ? span0 : span1;
      for (uint32_t off = 0; off < span; off += 32) {
        uint32_t b0[2], b1[2], rs[2], mine[2];
        bool valid[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t v = lo[q] + off + lane;
          valid[q] = v < hi[q];
          b0[q] = b1[q] = rs[q] = 0u; mine[q] = 0u;
          if (valid[q]) {
            b0[q] = s_bb[v]; b1[q] = s_bb[v + 1];
            const int32_t lim = (int32_t)keys[v] + W;
            uint32_t a = v, e = hi[q];  // last bin in [v, hi) with key <= lim
            while (e - a > 1) { const uint32_t mid = (a + e) >> 1; if ((int32_t)keys[mid] <= lim) a = mid; else e = mid; }
            const uint32_t reach = s_bb[a + 1], cap = b1[q] - 1 + L[q];
            rs[q] = reach < cap ? reach : cap;
          }
        }
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t B0 = __shfl_sync(0xFFFFFFFFu, b0[q], l), B1 = __shfl_sync(0xFFFFFFFFu, b1[q], l);
            const uint32_t RS = __shfl_sync(0xFFFFFFFFu, rs[q], l);
            const uint32_t cur = pos[q] > B0 ? pos[q] : B0;
            const uint32_t a = RS - cur;                      // meaningful when cur < B1 (then RS > cur)
            uint32_t rem = a - __umulhi(a, Mrec[q]) * L[q];   // a mod L, or a mod L + L
            rem = rem < rem - L[q] ? rem : rem - L[q];        // unsigned: picks the one below L
            const uint32_t p2 = RS - rem;                     // next unconsumed position after this bin's lobbies
            const bool inside = cur < B1, full = p2 >= B1;
            const uint32_t m = (!inside || full) ? B1 - B0 : p2 - B0;  // matched players of the bin (a prefix)
            pos[q] = !inside ? pos[q] : (full ? p2 : B1);     // the rest of a partly matched bin stays queued
            if ((int)lane == l) mine[q] = m;
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)
          if (valid[q]) s_m[lo[q] + off + lane] = mine[q];
      }
    }
    __syncthreads();
    if (!m_smem) {  // very large key domain: the counts were parked in global memory; scan them in place of the bases
      for (uint32_t v = tid; v < Kp; v += kScanBlock) s_bb[v] = s_m[v];
      s_m = s_bb;
      __syncthreads();
    }
    // member slots = exclusive scan of the matched counts (members of successive partitions are contiguous)
    n_matched = block_excl_scan<kScanBlock>(s_m, Kp, s_tmp);
    if (tid == 0) s_m[Kp] = n_matched;
    __syncthreads();
    for (uint32_t v = tid; v < Kp; v += kScanBlock) {
      t.outbase[v] = s_m[v];
      t.binlim[v] = s_m[v + 1];  // = outbase + matched players of the bin
    }
    for (uint32_t sg = tid; sg < n_segs; sg += kScanBlock) {
      const uint32_t mb = s_m[s_lo[sg]], nl = (s_m[s_lo[sg + 1]] - mb) / t.seg_L[sg];
      s_lob[sg] = nl;
      t.seg[sg].n_lobbies = nl; t.seg[sg].member_base = mb;
    }
    __syncthreads();
    tot_lob = block_excl_scan<kScanBlock>(s_lob, n_segs, s_tmp);
    for (uint32_t sg = tid; sg < n_segs; sg += kScanBlock) t.seg[sg].lobby_base = s_lob[sg];
  }
  if (tid == 0) {
    t.ctr->n_lobbies = tot_lob; t.ctr->n_matched = n_matched; t.ctr->n_alive = alive; t.ctr->n_dead = dead;
    // expected players of the fullest bin per tile of one row (players spread evenly over rows)
    const uint64_t npool = (uint64_t)alive + dead;
    t.ctr->heavy = ((uint64_t)s_misc[0] * kTile > 4ull * (npool ? npool : 1)) ? 1u : 0u;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kScanBlock) k_colscan(uint32_t R, uint32_t* __restrict__ M, const TailArgs t) {
  extern __shared__ __align__(16) uint32_t scratch[];  // max(kColScratchWords, tail_words(Kp, layout)) words
  if (blockIdx.x + 1 < gridDim.x) colscan_cols_body(scratch, blockIdx.x, R, t.Kp, M);
  else colscan_tail_body(scratch, t);
}

// ---------------------------------------------------------------------------------------
// k_place: the dominant kernel.  Row r walks its chunk in rounds of kRound players.
// For each player it needs the STABLE rank among the row's players of the same bin
// (deterministic tie-break by enqueue order).  Running slot counters cnt[bin] live in
// shared memory; inside a round:
//   S1  __match_any_sync groups a warp-batch by bin; each group's leader snapshots
//       cnt[bin] and pushes a node {prev, group size} on the bin's round-local list
//       (atomicExch on head[bin], epoch-tagged so stale heads read as empty).
//   S2  after a barrier each leader walks its bin's list: groups with a smaller node id
//       come earlier in enqueue order (node id = batch*kBlock + tid), so
//       slot = snapshot + sum(sizes of earlier groups) + rank inside the group.
//       The first pusher advances cnt[bin] by the round's total.
// Bit 31 of cnt marks a (row, bin) cell that reaches past the bin's matched prefix:
// only those players consult binlim (the leftovers of a partition stay queued, marked in left_bits).
// This is the round's first placement kernel, kept as an on-device cross-check of k_place2
// (rank_impl 1 = this list ranking, rank_impl 0 = a slow warp-serial ranking).
// ---------------------------------------------------------------------------------------
template <int IMPL>
__global__ void __launch_bounds__(kBlock, 1)
    k_place(PoolView p, uint32_t n, uint32_t chunk, BinMap bm, uint32_t Kp, uint32_t R, const uint32_t* __restrict__ M,
            const uint32_t* __restrict__ tot, const uint32_t* __restrict__ outbase,
            const uint32_t* __restrict__ binlim, uint64_t* __restrict__ members, uint32_t* __restrict__ src_idx,
            uint32_t* __restrict__ left_bits, uint32_t* __restrict__ rescnt, TickCtr* ctr) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* cnt = smem;
  uint32_t* head = cnt + Kp;                                 // IMPL 1 only
  uint32_t* node = head + (IMPL == 1 ? Kp : 0);              // [kRound]
  uint16_t* s_lut = reinterpret_cast<uint16_t*>(node + (IMPL == 1 ? kRound : 0));
  __shared__ uint32_t s_nres;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t row = blockIdx.x;
  const uint64_t pol_in = policy_evict_first(), pol_out = policy_evict_last();
  {
    const uint32_t* mrow = M + (size_t)row * Kp;
    const uint32_t* mnext = (row + 1 < R) ? mrow + Kp : tot;  // prefix of the next row, or column total
    for (uint32_t i = tid; i < Kp; i += kBlock) {
      uint32_t v = 0;
      if (i < bm.K) {
        const uint32_t pre = mrow[i], c = mnext[i] - pre;
        const uint32_t start = outbase[i] + pre;  // slot of the cell's first player
        v = start | ((start + c > binlim[i]) ? 0x80000000u : 0u);
      }
      cnt[i] = v;
      if (IMPL == 1) head[i] = 0;
    }
    for (uint32_t i = tid; i < bm.KR; i += kBlock) s_lut[i] = bm.lut[i];
    if (tid == 0) s_nres = 0;
  }
  __syncthreads();

  const uint64_t beg64 = (uint64_t)row * chunk;
  const uint32_t beg = beg64 < n ? (uint32_t)beg64 : n;
  const uint32_t end = (beg64 + chunk < n) ? (uint32_t)(beg64 + chunk) : n;
  const uint32_t n_rounds = (end - beg + kRound - 1) / kRound;

  uint32_t nleft = 0;
  for (uint32_t round = 0; round < n_rounds; ++round) {
    const uint32_t base = beg + round * kRound;
    uint32_t bin[kJ];
    uint64_t idv[kJ];
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const uint32_t e = base + j * kBlock + tid;
      if (e < end) {
        int32_t r; uint32_t m;
        r = ld_stream_s32(p.rating + e, pol_in);
        m = ld_stream_u8(p.mode + e, pol_in);
        idv[j] = ld_stream_u64(p.id + e, pol_in);
        bin[j] = bin_of(bm, s_lut, r, m);
      } else {
        bin[j] = 0xFFFFFFFFu;
        idv[j] = 0;
      }
    }
    uint32_t leader[kJ], rankw[kJ], base_g[kJ];

    if (IMPL == 1) {
      const uint32_t epoch = round + 1;
      uint32_t snap[kJ], mynode[kJ];
      bool isl[kJ], first[kJ];
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const uint32_t mask = __match_any_sync(0xFFFFFFFFu, bin[j]);
        leader[j] = __ffs(mask) - 1;
        rankw[j] = __popc(mask & lt_mask);
        isl[j] = (lane == leader[j]) && (bin[j] < bm.K);
        first[j] = false;
        mynode[j] = j * kBlock + tid;
        snap[j] = 0;
        if (isl[j]) {
          snap[j] = cnt[bin[j]];
          const uint32_t prev = atomicExch(&head[bin[j]], (epoch << 13) | mynode[j]);
          const uint32_t prevnode = ((prev >> 13) == epoch) ? (prev & kNone) : kNone;
          node[mynode[j]] = prevnode | ((uint32_t)__popc(mask) << 13);
          first[j] = (prevnode == kNone);
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        uint32_t bg = 0;
        if (isl[j]) {
          uint32_t cur = head[bin[j]] & kNone, lower = 0, total = 0;
          while (cur != kNone) {
            const uint32_t nd = node[cur];
            const uint32_t c = nd >> 13;
            total += c;
            if (cur < mynode[j]) lower += c;
            cur = nd & kNone;
          }
          bg = snap[j] + lower;
          if (first[j]) cnt[bin[j]] = snap[j] + total;
        }
        base_g[j] = __shfl_sync(0xFFFFFFFFu, bg, leader[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        base_g[j] = 0;
        for (uint32_t w = 0; w < kBlock / 32; ++w) {
          if (warp == w) {
            const uint32_t mask = __match_any_sync(0xFFFFFFFFu, bin[j]);
            leader[j] = __ffs(mask) - 1;
            rankw[j] = __popc(mask & lt_mask);
            uint32_t bg = 0;
            if (lane == leader[j] && bin[j] < bm.K) {
              bg = cnt[bin[j]];
              cnt[bin[j]] = bg + __popc(mask);
            }
            base_g[j] = __shfl_sync(0xFFFFFFFFu, bg, leader[j]);
          }
          __syncthreads();
        }
      }
    }

#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      bool left = false;
      if (bin[j] < bm.K) {
        const uint32_t e = base + j * kBlock + tid;
        uint32_t slot = (base_g[j] & 0x7FFFFFFFu) + rankw[j];
        bool matched = true;
        if (base_g[j] >> 31) matched = slot < __ldg(&binlim[bin[j]]);
        if (matched) {
          st_keep_u64(members + slot, idv[j], pol_out);
          if (src_idx) src_idx[slot] = e;
        } else {
          left = true;
        }
      }
      // one bit per player that stays queued; the warp's 32 positions of a batch are one word
      const uint32_t wv = __ballot_sync(0xFFFFFFFFu, left);
      if (lane == 0) { left_bits[(base + j * kBlock + warp * 32) >> 5] = wv; nleft += __popc(wv); }
    }
    if (IMPL == 1) __syncthreads();
  }
  if (lane == 0 && nleft) atomicAdd(&s_nres, nleft);
  __syncthreads();
  if (tid == 0) rescnt[row] = s_nres;
}

// ---------------------------------------------------------------------------------------
// k_place2<BLOCK>: the production placement kernel (rank_impl 3).  Same contract as
// k_place, but
//   * the row's (bin u16, id u64) columns arrive as 2 048-player tiles through a ring of
//     TMA bulk copies (cp.async.bulk -> mbarrier), issued `stages` tiles ahead by one
//     thread: DRAM latency never stalls the ranking, inputs stream with L2 evict-first;
//   * light bins (the normal case with ~5k rating values per mode): ONE list node per
//     player, no warp vote — push on a HASHED head table (kHeadSlots entries, epoch-tagged,
//     never cleared) with a shared-memory atomicExch, barrier, walk the slot's round-local
//     list counting same-bin nodes with a smaller tile position; the lowest one advances
//     the bin's slot counter.  Per-CTA state is 4 B/bin + 28 KB, so two CTAs share an SM
//     and one CTA's barrier phases overlap the other's work;
//   * heavy bins (k_colscan flags the tick when some bin expects > 4 players per tile,
//     e.g. everyone at the default rating): warp-aggregated groups, lists <= 64 nodes;
//   * few bins (arrival order: bin = (mode, group), <= 256): dense per-(bin, warp-batch)
//     group-size matrix + one warp-shuffle scan per bin;
//   * ids are stored with an L2 evict-last policy: the 4 writes completing a 32-byte
//     sector of member_ids arrive at unrelated times and must meet in L2, not in DRAM.
// Shared memory: ring | mbarriers | cnt[Kp] | head[kHeadSlots] | node[kTile] | nbin | dense.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kHeadSlots = 4096;

template <int BLOCK>
__device__ __forceinline__ void place2_body(
    unsigned char* smem_raw, const uint16_t* __restrict__ bins16, const uint64_t* __restrict__ ids, uint32_t n,
    uint32_t chunk, uint32_t K, uint32_t Kp, uint32_t R, uint32_t stages, uint32_t dense, const uint32_t* __restrict__ M,
    const uint32_t* __restrict__ tot, const uint32_t* __restrict__ outbase, const uint32_t* __restrict__ binlim,
    uint64_t* __restrict__ members, uint32_t* __restrict__ src_idx, uint32_t* __restrict__ left_bits,
    uint32_t* __restrict__ rescnt, TickCtr* ctr, uint32_t dbg_all) {
  const uint32_t dbg = dbg_all & 3u;  // (higher bits are histogram-phase experiments)
  // dbg != 0: timing experiments only (results invalid): 1 = rank, no id store; 2 = no rank,
  // coalesced store; 3 = no rank, pseudo-random scatter
  constexpr int J = kTile / BLOCK;
  constexpr int NW = BLOCK / 32;
  uint64_t* ring_ids = reinterpret_cast<uint64_t*>(smem_raw);                               // [stages][kTile]
  uint16_t* ring_bins = reinterpret_cast<uint16_t*>(smem_raw + (size_t)stages * kTile * 8);  // [stages][kTile]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * kTileBytes);      // [kMaxStages]
  uint32_t* cnt = reinterpret_cast<uint32_t*>(smem_raw + (size_t)stages * kTileBytes + 64);  // [Kp]
  uint32_t* head = cnt + Kp;                 // [kHeadSlots]
  uint32_t* node = head + kHeadSlots;        // [kTile]
  uint16_t* nbin = reinterpret_cast<uint16_t*>(node + kTile);  // [kTile] heavy path: bin of a group node
  uint16_t* wc = nbin + kTile;                                     // dense only: [Kp][kDenseStride] group sizes
  uint16_t* pf = wc + (size_t)Kp * kDenseStride;                    // dense only: their exclusive prefixes
  uint32_t* cbase = reinterpret_cast<uint32_t*>(pf + (size_t)Kp * kDenseStride);  // dense only: [Kp]
  __shared__ uint32_t s_nres;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t row = blockIdx.x;
  const uint64_t pol_in = policy_evict_first(), pol_out = policy_evict_last();

  const uint64_t beg64 = (uint64_t)row * chunk;
  const uint32_t beg = beg64 < n ? (uint32_t)beg64 : n;
  const uint32_t end = (beg64 + chunk < n) ? (uint32_t)(beg64 + chunk) : n;
  const uint32_t n_tiles = (end - beg + kTile - 1) / kTile;

  if (tid == 0) {
    for (uint32_t s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
    s_nres = 0;
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {  // prologue: fill the ring (whole tiles; the pool columns are padded past n)
    for (uint32_t t = 0; t < stages && t < n_tiles; ++t) {
      mbar_expect_tx(&full[t], kTileBytes);
      tma_load_1d(ring_ids + (size_t)t * kTile, ids + beg + (size_t)t * kTile, kTile * 8, &full[t], pol_in);
      tma_load_1d(ring_bins + (size_t)t * kTile, bins16 + beg + (size_t)t * kTile, kTile * 2, &full[t], pol_in);
    }
  }
  {
    const uint32_t* mrow = M + (size_t)row * Kp;
    const uint32_t* mnext = (row + 1 < R) ? mrow + Kp : tot;
    for (uint32_t i = tid; i < Kp; i += BLOCK) {
      uint32_t v = 0;
      if (i < K) {  // __ldcg: these arrays are produced earlier in the same (fused) launch by other SMs
        const uint32_t pre = __ldcg(&mrow[i]), c = __ldcg(&mnext[i]) - pre;
        const uint32_t start = __ldcg(&outbase[i]) + pre;  // slot of the cell's first player
        v = start | ((start + c > __ldcg(&binlim[i])) ? 0x80000000u : 0u);
      }
      cnt[i] = v;
    }
    for (uint32_t i = tid; i < kHeadSlots; i += BLOCK) head[i] = 0;
    if (dense)
      for (uint32_t i = tid; i < Kp * kDenseStride / 2; i += BLOCK) reinterpret_cast<uint32_t*>(wc)[i] = 0;
  }
  const bool heavy = __ldcg(&ctr->heavy) != 0;
  __syncthreads();

  uint32_t st = 0, parity = 0;
  uint32_t nleft = 0;  // lane 0: players of this warp's positions that stay queued
  for (uint32_t t = 0; t < n_tiles; ++t) {
    const uint32_t tile_base = beg + t * kTile;
    const uint32_t valid = end - tile_base;  // players of this tile inside the row (>= kTile except the last)
    const uint16_t* tb = ring_bins + (size_t)st * kTile;
    const uint64_t* ti = ring_ids + (size_t)st * kTile;
    mbar_wait(&full[st], parity);
    const uint32_t epoch = t + 1;
    uint32_t bin[J], slot[J], pos_[J];
    uint64_t idv[J];
    bool flag[J];
    if (J == 4 && dense == 2) {
      // blocked arrangement (thread t owns 4 consecutive tile positions): one 64-bit load of the
      // 4 bins, two 128-bit loads of the 4 ids — strided scalar loads would be 8-way bank conflicts
      const uint2 bb = reinterpret_cast<const uint2*>(tb)[tid];
      const uint4 i01 = reinterpret_cast<const uint4*>(ti)[2 * tid], i23 = reinterpret_cast<const uint4*>(ti)[2 * tid + 1];
      const uint32_t b4[4] = {bb.x & 0xFFFFu, bb.x >> 16, bb.y & 0xFFFFu, bb.y >> 16};
      const uint64_t i4[4] = {(uint64_t)i01.x | ((uint64_t)i01.y << 32), (uint64_t)i01.z | ((uint64_t)i01.w << 32),
                              (uint64_t)i23.x | ((uint64_t)i23.y << 32), (uint64_t)i23.z | ((uint64_t)i23.w << 32)};
#pragma unroll
      for (int j = 0; j < J; ++j) {
        pos_[j] = tid * J + j;
        bin[j] = (pos_[j] < valid) ? b4[j & 3] : 0xFFFFu;
        idv[j] = i4[j & 3];
        slot[j] = 0; flag[j] = false;
      }
    } else {
#pragma unroll
      for (int j = 0; j < J; ++j) {
        pos_[j] = j * BLOCK + tid;
        bin[j] = (pos_[j] < valid) ? (uint32_t)tb[pos_[j]] : 0xFFFFu;
        slot[j] = 0; flag[j] = false;
      }
    }
    if (dbg >= 2) {
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const uint32_t e = tile_base + j * BLOCK + tid;
        slot[j] = dbg == 2 ? e : (uint32_t)(((uint64_t)e * 2654435761ull) % n);
        idv[j] = ti[j * BLOCK + tid];
      }
    } else if (dense == 2) {
      // Few bins, no warp vote (MATCH.ANY costs 64 cycles per warp instruction per SM on B200):
      // every thread counts its own J consecutive players in private byte counters
      // c8[bin][thread], one warp-shuffle scan per bin turns them into per-16-thread bases,
      // and a thread's offset inside its 16-group is a masked byte sum (dp4a).
      uint8_t* c8 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wc) + 15) & ~uintptr_t(15));  // [Kp][BLOCK]
      // rows padded (+16 B, +2 entries) so that different bins fall into different banks
      constexpr uint32_t CS = BLOCK + 16, LS = 34;
      uint16_t* lb = reinterpret_cast<uint16_t*>(c8 + (size_t)Kp * CS);               // [Kp][LS]
      uint32_t* cb2 = reinterpret_cast<uint32_t*>(lb + (size_t)Kp * LS);              // [Kp]
      for (uint32_t i = tid; i < Kp * (CS / 16); i += BLOCK) reinterpret_cast<uint4*>(c8)[i] = make_uint4(0, 0, 0, 0);
      __syncthreads();
      uint32_t lrank[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        lrank[j] = 0;
        if (bin[j] < K) {
          uint8_t* c = c8 + bin[j] * CS + tid;
          lrank[j] = *c;
          *c = (uint8_t)(lrank[j] + 1);
        }
      }
      __syncthreads();
      for (uint32_t b = warp; b < K; b += NW) {  // lane l sums the counters of threads 16l .. 16l+15
        const uint4 v = reinterpret_cast<const uint4*>(c8 + b * CS)[lane];
        uint32_t incl = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
        const uint32_t own = incl;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, off);
          if (lane >= (uint32_t)off) incl += u;
        }
        lb[b * LS + lane] = (uint16_t)(incl - own);
        if (lane == 31) { const uint32_t base = cnt[b]; cb2[b] = base; cnt[b] = base + incl; }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (bin[j] < K) {
          const uint32_t g = tid >> 4, k = tid & 15;
          const uint4 v = reinterpret_cast<const uint4*>(c8 + bin[j] * CS)[g];
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
          uint32_t pre = 0;
#pragma unroll
          for (int wi = 0; wi < 4; ++wi) {
            const uint32_t m = ((uint32_t)wi < (k >> 2)) ? 0x01010101u
                               : ((uint32_t)wi == (k >> 2) ? (((1u << (8 * (k & 3))) - 1u) & 0x01010101u) : 0u);
            pre = __dp4a(w[wi], m, pre);
          }
          const uint32_t base = cb2[bin[j]];
          slot[j] = (base & 0x7FFFFFFFu) + lb[bin[j] * LS + g] + pre + lrank[j];
          flag[j] = (base >> 31) != 0;
        }
      }
    } else if (dense) {
      // per-(bin, warp-batch) group sizes in a small matrix, one shuffle scan per bin across
      // the tile's 64 warp-batches (batch = j * NW + warp, increasing with tile position)
      uint32_t rankw[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const uint32_t mask = __match_any_sync(0xFFFFFFFFu, bin[j]);
        rankw[j] = __popc(mask & lt_mask);
        if (lane == (uint32_t)(__ffs(mask) - 1) && bin[j] < K)
          wc[bin[j] * kDenseStride + j * NW + warp] = (uint16_t)__popc(mask);
      }
      __syncthreads();
      for (uint32_t b = warp; b < K; b += NW) {  // lane l owns warp-batches 2l, 2l+1
        uint32_t* w32 = reinterpret_cast<uint32_t*>(wc + b * kDenseStride) + lane;
        const uint32_t two = *w32;
        *w32 = 0;  // the matrix is all-zero again for the next tile
        const uint32_t c0 = two & 0xFFFFu, c1 = two >> 16;
        uint32_t incl = c0 + c1;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
          if (lane >= (uint32_t)off) incl += v;
        }
        const uint32_t excl = incl - c0 - c1;
        reinterpret_cast<uint32_t*>(pf + b * kDenseStride)[lane] = excl | ((excl + c0) << 16);
        if (lane == 31) { const uint32_t base = cnt[b]; cbase[b] = base; cnt[b] = base + incl; }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (bin[j] < K) {
          const uint32_t base = cbase[bin[j]];
          slot[j] = (base & 0x7FFFFFFFu) + pf[bin[j] * kDenseStride + j * NW + warp] + rankw[j];
          flag[j] = (base >> 31) != 0;
        }
      }
    } else if (!heavy) {
      uint32_t snap[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const uint32_t pos = j * BLOCK + tid;
        snap[j] = 0;
        if (bin[j] < K) {
          snap[j] = cnt[bin[j]];
          const uint32_t prev = atomicExch(&head[bin[j] & (kHeadSlots - 1)], (epoch << 12) | pos);
          const uint32_t pn = ((prev >> 12) == epoch) ? (prev & 0xFFFu) : 0xFFFu;
          node[pos] = pn | (bin[j] << 12);
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < J; ++j) idv[j] = ti[j * BLOCK + tid];  // ids early: their latency hides behind the walks
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const uint32_t pos = j * BLOCK + tid;
        if (bin[j] < K) {
          uint32_t cur = head[bin[j] & (kHeadSlots - 1)] & 0xFFFu, lower = 0, total = 0;
          while (cur != 0xFFFu) {
            const uint32_t nd = node[cur];
            if ((nd >> 12) == bin[j]) {  // the slot is shared by bins congruent mod kHeadSlots
              ++total;
              lower += (cur < pos) ? 1u : 0u;
            }
            cur = nd & 0xFFFu;
          }
          slot[j] = (snap[j] & 0x7FFFFFFFu) + lower;
          flag[j] = (snap[j] >> 31) != 0;
          if (lower == 0) cnt[bin[j]] = snap[j] + total;  // the bin's earliest player of the tile
        }
      }
    } else {
      uint32_t snap[J], leader[J], rankw[J];
      bool isl[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const uint32_t pos = j * BLOCK + tid;
        const uint32_t mask = __match_any_sync(0xFFFFFFFFu, bin[j]);
        leader[j] = __ffs(mask) - 1;
        rankw[j] = __popc(mask & lt_mask);
        isl[j] = (lane == leader[j]) && (bin[j] < K);
        snap[j] = 0;
        if (isl[j]) {
          snap[j] = cnt[bin[j]];
          const uint32_t prev = atomicExch(&head[bin[j] & (kHeadSlots - 1)], (epoch << 12) | pos);
          const uint32_t pn = ((prev >> 12) == epoch) ? (prev & 0xFFFu) : 0xFFFu;
          node[pos] = pn | ((uint32_t)__popc(mask) << 12);
          nbin[pos] = (uint16_t)bin[j];
        }
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const uint32_t pos = j * BLOCK + tid;
        uint32_t bg = 0;
        if (isl[j]) {
          uint32_t cur = head[bin[j] & (kHeadSlots - 1)] & 0xFFFu, lower = 0, total = 0;
          while (cur != 0xFFFu) {
            const uint32_t nd = node[cur];
            if (nbin[cur] == bin[j]) {
              const uint32_t c = nd >> 12;
              total += c;
              if (cur < pos) lower += c;
            }
            cur = nd & 0xFFFu;
          }
          bg = snap[j] + lower;
          if (lower == 0) cnt[bin[j]] = snap[j] + total;
        }
        bg = __shfl_sync(0xFFFFFFFFu, bg, leader[j]);
        slot[j] = (bg & 0x7FFFFFFFu) + rankw[j];
        flag[j] = (bg >> 31) != 0;
      }
    }
    if (dbg < 2 && (dense == 1 || (!dense && heavy))) {
#pragma unroll
      for (int j = 0; j < J; ++j) idv[j] = ti[pos_[j]];
    }
    // ---- store matched ids; players past their bin's matched prefix stay queued: one bit per player in
    // left_bits (every word of the row is written every tick, no atomics, no cold branch in this loop) ----
    uint32_t lmask = 0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      if (bin[j] < K) {
        bool matched = true;
        if (flag[j]) matched = slot[j] < __ldcg(&binlim[bin[j]]);
        if (dbg == 1) continue;
        if (matched) {
          st_keep_u64(members + slot[j], idv[j], pol_out);
          if (src_idx) src_idx[slot[j]] = tile_base + pos_[j];
        } else {
          lmask |= 1u << j;
        }
      }
    }
    {
      uint32_t* lw = left_bits + ((tile_base + warp * (32 * J)) >> 5);
      if (J == 4 && dense == 2) {  // blocked: the warp owns 128 consecutive positions, lane l the bits 4l .. 4l+3
        const uint32_t mine = lmask << ((lane & 7u) * 4u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t wv = __reduce_or_sync(0xFFFFFFFFu, (lane >> 3) == (uint32_t)k ? mine : 0u);
          if (lane == 0) { lw[k] = wv; nleft += __popc(wv); }
        }
      } else {                     // strided: batch j of the warp = positions j*BLOCK + 32*warp .. +31 = one word
        uint32_t mine = 0, all = 0;    // lane j stores batch j's word: one store instruction per warp and tile
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const uint32_t wv = __ballot_sync(0xFFFFFFFFu, (lmask >> j) & 1u);
          if (lane == (uint32_t)j) mine = wv;
          all += __popc(wv);
        }
        if (lane < (uint32_t)J) left_bits[(tile_base + lane * BLOCK + warp * 32) >> 5] = mine;
        if (lane == 0) nleft += all;
      }
    }
    __syncthreads();  // everyone is done with stage st and with this round's lists
    if (tid == 0 && t + stages < n_tiles) {
      const uint32_t tn = t + stages;
      mbar_expect_tx(&full[st], kTileBytes);
      tma_load_1d(ring_ids + (size_t)st * kTile, ids + beg + (size_t)tn * kTile, kTile * 8, &full[st], pol_in);
      tma_load_1d(ring_bins + (size_t)st * kTile, bins16 + beg + (size_t)tn * kTile, kTile * 2, &full[st], pol_in);
    }
    if (++st == stages) { st = 0; parity ^= 1u; }
  }

  if (lane == 0 && nleft) atomicAdd(&s_nres, nleft);
  __syncthreads();
  if (tid == 0) rescnt[row] = s_nres;  // players of this row that stay queued
  if (tid == 0)
    for (uint32_t s = 0; s < stages; ++s) mbar_inval(&full[s]);
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, (BLOCK == 512 ? 2 : 1))
    k_place2(const uint16_t* __restrict__ bins16, const uint64_t* __restrict__ ids, uint32_t n, uint32_t chunk, uint32_t K,
             uint32_t Kp, uint32_t R, uint32_t stages, uint32_t dense, const uint32_t* __restrict__ M,
             const uint32_t* __restrict__ tot, const uint32_t* __restrict__ outbase,
             const uint32_t* __restrict__ binlim, uint64_t* __restrict__ members, uint32_t* __restrict__ src_idx,
             uint32_t* __restrict__ left_bits, uint32_t* __restrict__ rescnt, TickCtr* ctr, uint32_t dbg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  place2_body<BLOCK>(smem_raw, bins16, ids, n, chunk, K, Kp, R, stages, dense, M, tot, outbase, binlim, members, src_idx,
                     left_bits, rescnt, ctr, dbg);
}

// ---------------------------------------------------------------------------------------
// k_epilogue.  Lobby headers from the segment table — lobby c of segment s = members
// [member_base + k*L, +L); replaces the payload assembly at search/worker.ex:315-319.
// Pool compaction, row-parallel and order-preserving: the placement pass left one bit per
// player that stays queued (left_bits) and the count per row; every CTA scans the R row
// counts, then walks the bit words of its rows — popcount prefix, slots of the set bits
// enumerated into shared memory, one thread per leftover player gathers its record from the
// old pool buffer into the alternate one and re-stamps the player's active-set entry.
// Replaces save_new_state/3 (search/worker.ex:282-289): the "partial lobby" is the players
// left resident.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kLeftList = 2048;  // leftover players handled per step of the compaction
constexpr uint32_t kEpiScratchWords = (kMaxRows + 1) + 64 + (kMaxSegs + 1) + 2 * kMaxSegs + kLeftList;

template <int BLOCK>
__device__ __forceinline__ void epilogue_body(uint32_t* scratch, PoolView src, PoolView dst, uint32_t n, uint32_t chunk,
                                              uint32_t R, const uint32_t* __restrict__ rescnt,
                                              const uint32_t* __restrict__ left_bits, ActiveView act, uint32_t new_gen,
                                              const SegInfo* __restrict__ seg, const uint32_t* __restrict__ seg_L,
                                              uint32_t n_segs, uint32_t n_groups, mm_lobby_hdr* __restrict__ hdr,
                                              const uint32_t* __restrict__ src_idx, uint32_t* __restrict__ emit_seq,
                                              uint32_t* __restrict__ tot, uint32_t Kp, TickCtr* ctr,
                                              unsigned long long* t_mid = nullptr) {
  constexpr uint32_t NW = BLOCK / 32;
  uint32_t* s_off = scratch;                   // [kMaxRows + 1]
  uint32_t* s_tmp = s_off + kMaxRows + 1;      // [64]
  uint32_t* s_lbase = s_tmp + 64;              // [kMaxSegs + 1]
  uint32_t* s_mbase = s_lbase + kMaxSegs + 1;  // [kMaxSegs]
  uint32_t* s_L = s_mbase + kMaxSegs;          // [kMaxSegs]
  uint32_t* s_list = s_L + kMaxSegs;           // [kLeftList]
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (uint32_t s = tid; s < n_segs; s += BLOCK) {
    s_lbase[s] = __ldcg(&seg[s].lobby_base); s_mbase[s] = __ldcg(&seg[s].member_base); s_L[s] = seg_L[s];
  }
  for (uint32_t r = tid; r < R; r += BLOCK) s_off[r] = __ldcg(&rescnt[r]);
  __syncthreads();
  for (uint32_t i = blockIdx.x * BLOCK + tid; i < Kp; i += gridDim.x * BLOCK) tot[i] = 0;  // ready for the next tick
  const uint32_t total = block_excl_scan<BLOCK>(s_off, R, s_tmp);
  if (tid == 0) {
    s_off[R] = total;
    if (blockIdx.x == 0) ctr->n_resid = total;
  }
  __syncthreads();
  // Work is split by leftover RANK, not by row: under policy S0 the leftovers are the latest arrivals of every
  // partition and sit in the last rows of the pool.  CTA b moves the players with global rank [r0, r1); it walks
  // the bit words of the rows holding them (popcount prefix from the start of the row), enumerates the pool
  // slots of its ranks into a shared-memory list (no memory latency) and then, one thread per listed player,
  // gathers the record into the alternate pool buffer and re-stamps the player's active-set entry — all the
  // dependent gather / hash-probe chains run in parallel, neighbouring threads touch neighbouring slots.
  const uint32_t per = (total + gridDim.x - 1) / gridDim.x;
  const uint32_t r0 = (uint64_t)blockIdx.x * per < total ? blockIdx.x * per : total;
  const uint32_t r1 = r0 + per < total ? r0 + per : total;
  if (r1 > r0) {
    uint32_t tbase = r0, fill = 0;  // global rank of s_list[0]; entries in the list (uniform)
    uint32_t row_beg = 0;           // pool slot of the current row's first player
    auto flush = [&](uint32_t count, bool last) {
      __syncthreads();
      for (uint32_t e = tid; e < count; e += BLOCK) {
        const uint32_t i = s_list[e], t = tbase + e;
        const uint64_t pid = src.id[i];
        dst.id[t] = pid; dst.rating[t] = src.rating[i]; dst.mode[t] = src.mode[i];
        dst.tsize[t] = src.tsize[i]; dst.ts[t] = src.ts[i]; dst.bin[t] = src.bin[i];
        if (act.mask) {
          uint64_t h = hash64(pid) & act.mask;
          for (uint64_t probe = 0; probe <= act.mask; ++probe) {
            const unsigned long long k2 = act.keys[h];
            if (k2 == pid) { act.vals[h] = ((unsigned long long)new_gen << 32) | t; break; }
            if (k2 == kEmptyKey) break;
            h = (h + 1) & act.mask;
          }
        }
      }
      tbase += count;
      if (!last) __syncthreads();  // the last flush runs on into the lobby headers: the few threads waiting on
                                   // their gather / probe chains do not hold up the others
    };
    uint32_t row = 0;
    {  // first row holding rank r0: smallest row with s_off[row + 1] > r0
      uint32_t a = 0, e = R;
      while (a < e) { const uint32_t mid = (a + e) >> 1; if (s_off[mid + 1] > r0) e = mid; else a = mid + 1; }
      row = a;
    }
    for (; row < R && s_off[row] < r1; ++row) {
      const uint32_t off = s_off[row], cnt = s_off[row + 1] - off;
      if (cnt == 0) continue;  // uniform for the CTA
      const uint64_t beg64 = (uint64_t)row * chunk;
      const uint32_t beg = beg64 < n ? (uint32_t)beg64 : n;
      const uint32_t end = (beg64 + chunk < n) ? (uint32_t)(beg64 + chunk) : n;
      const uint32_t nwords = (end - beg + 31) >> 5;  // beg is a multiple of 32 (chunk is a multiple of kRound)
      const uint32_t* bits = left_bits + (beg >> 5);
      row_beg = beg;
      const uint32_t lo_l = (r0 > off ? r0 : off) - off, hi_l = (r1 < off + cnt ? r1 : off + cnt) - off;  // row-local ranks
      uint32_t run_l = 0;  // row-local rank of the step's first leftover player
      for (uint32_t w0 = 0; w0 < nwords && run_l < hi_l; w0 += BLOCK) {  // BLOCK words = 32 * BLOCK players per step
        const uint32_t wi = w0 + tid;
        const uint32_t w = wi < nwords ? __ldcg(&bits[wi]) : 0u;
        const uint32_t c = __popc(w);
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
          if (lane >= (uint32_t)o) incl += u;
        }
        if (lane == 31) s_tmp[warp] = incl;
        __syncthreads();
        uint32_t wbase = 0, wtot = 0;
        for (uint32_t k = 0; k < NW; ++k) { const uint32_t v = s_tmp[k]; if (k < warp) wbase += v; wtot += v; }
        const uint32_t lpre = run_l + wbase + incl - c;  // row-local rank of this word's first leftover player
        uint32_t q = lo_l > run_l ? lo_l : run_l;
        const uint32_t q_end = hi_l < run_l + wtot ? hi_l : run_l + wtot;
        while (q < q_end) {  // (uniform) ranks [q, q_end) of this step are mine
          if (fill == kLeftList) { flush(fill, false); fill = 0; }
          const uint32_t room = kLeftList - fill, take = q_end - q < room ? q_end - q : room;
          if (c && lpre < q + take && lpre + c > q) {
            uint32_t ww = w, r = lpre;
            while (ww) {
              const uint32_t bpos = __ffs(ww) - 1;
              ww &= ww - 1;
              if (r >= q && r < q + take) s_list[fill + (r - q)] = row_beg + (wi << 5) + bpos;
              ++r;
            }
          }
          fill += take;
          q += take;
        }
        run_l += wtot;
        __syncthreads();  // s_tmp is rewritten by the next step
      }
    }
    if (fill) flush(fill, true);
  }
  if (t_mid && tid == 0) {
    unsigned long long tm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm));
    atomicMax(t_mid, tm);
  }
  const uint32_t total_lob = __ldcg(&ctr->n_lobbies);
  for (uint32_t c = blockIdx.x * BLOCK + tid; c < total_lob; c += gridDim.x * BLOCK) {
    uint32_t a = 0, e = n_segs;  // last segment with lobby_base <= c
    while (e - a > 1) { const uint32_t mid = (a + e) >> 1; if (s_lbase[mid] <= c) a = mid; else e = mid; }
    const uint32_t L = s_L[a];
    mm_lobby_hdr h;
    h.first_member = s_mbase[a] + (c - s_lbase[a]) * L;
    h.n_members = (uint16_t)L;
    h.mode = (uint8_t)(a / n_groups);
    h.group = (uint8_t)(a % n_groups);
    hdr[c] = h;
    if (emit_seq) emit_seq[c] = __ldcg(&src_idx[h.first_member + L - 1]);
  }
}

__global__ void __launch_bounds__(1024) k_epilogue(PoolView src, PoolView dst, uint32_t n, uint32_t chunk, uint32_t R,
                                                   const uint32_t* __restrict__ rescnt,
                                                   const uint32_t* __restrict__ left_bits, ActiveView act, uint32_t new_gen,
                                                   const SegInfo* __restrict__ seg, const uint32_t* __restrict__ seg_L,
                                                   uint32_t n_segs, uint32_t n_groups, mm_lobby_hdr* __restrict__ hdr,
                                                   const uint32_t* __restrict__ src_idx, uint32_t* __restrict__ emit_seq,
                                                   uint32_t* __restrict__ tot, uint32_t Kp, TickCtr* ctr) {
  __shared__ uint32_t scratch[kEpiScratchWords];
  epilogue_body<1024>(scratch, src, dst, n, chunk, R, rescnt, left_bits, act, new_gen, seg, seg_L, n_segs, n_groups, hdr,
                      src_idx, emit_seq, tot, Kp, ctr);
}

// ---------------------------------------------------------------------------------------
// k_tick<512>: the whole search tick in ONE cooperative launch ("fully matched in one
// launch", BASELINE.json).  Phases are the bodies above, separated by grid barriers; the
// CTA's dynamic shared memory is re-used by every phase:
//   hist (TMA ring of rating/mode tiles, row histogram, bin column)      | barrier 1
//   column scan of M (all CTAs) + tail (last CTA: bin bases, segments)    | barrier 2
//   placement (TMA ring of bin/id tiles, stable ranks, id scatter)        | barrier 3
//   epilogue (lobby headers by all CTAs, pool compaction by CTA 0)
// Saves three launch boundaries and their prologues (~10 us each on B200).
// ---------------------------------------------------------------------------------------
struct TickArgs {
  PoolView src, dst;
  uint32_t n, chunk, R, n_groups, hist_stages, place_stages, dense, new_gen, dbg;
  uint32_t* M;
  TailArgs tail;  // Kp, K, n_segs, tot, segment tables, outbase / binlim, counters
  uint32_t* tot;  // = tail.tot (written by the histogram and re-zeroed by the epilogue)
  uint64_t* members; uint32_t* src_idx; mm_lobby_hdr* hdr; uint32_t* emit_seq;
  uint32_t* left_bits;  // one bit per pool slot: the player stays queued after this tick
  uint32_t* rescnt; ActiveView act;
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 2) k_tick(const TickArgs a) {
  static_assert(BLOCK == kScanBlock, "the column-scan phase is written for 512-thread CTAs");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* scratch = reinterpret_cast<uint32_t*>(smem_raw);
  const unsigned int G = gridDim.x;
  const uint32_t Kp = a.tail.Kp, K = a.tail.K;
  TickCtr* ctr = a.tail.ctr;
  auto stamp = [&](int k) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      ctr->t[k] = t;
    }
  };
  stamp(0);
  hist3_body<BLOCK>(smem_raw, a.src.bin, a.n, a.chunk, Kp, a.hist_stages, a.M, a.tot);
  grid_barrier(&ctr->gbar, G);
  stamp(1);
  if (blockIdx.x == G - 1) {
    colscan_tail_body(scratch, a.tail);
    if (threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      ctr->t[5] = t;
    }
  }
  for (uint32_t g = blockIdx.x; g < (Kp + 31) / 32; g += G) colscan_cols_body(scratch, g, a.R, Kp, a.M);
  grid_barrier(&ctr->gbar, 2 * G);
  stamp(2);
  place2_body<BLOCK>(smem_raw, a.src.bin, a.src.id, a.n, a.chunk, K, Kp, a.R, a.place_stages, a.dense, a.M, a.tot,
                     a.tail.outbase, a.tail.binlim, a.members, a.src_idx, a.left_bits, a.rescnt, ctr, a.dbg);
  grid_barrier(&ctr->gbar, 3 * G);
  stamp(3);
  epilogue_body<BLOCK>(scratch, a.src, a.dst, a.n, a.chunk, a.R, a.rescnt, a.left_bits, a.act, a.new_gen, a.tail.seg,
                       a.tail.seg_L, a.tail.n_segs, a.n_groups, a.hdr, a.src_idx, a.emit_seq, a.tot, Kp, ctr, &ctr->t[7]);
  stamp(4);  // CTA 0's view
  if (threadIdx.x == 0) {  // the last CTA to finish closes the epilogue phase
    unsigned long long tm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm));
    atomicMax(&ctr->t[6], tm);
  }
}

// =======================================================================================
// Active set (replaces the Mnesia table of models/active_user.ex) + pool ingest.
// Open addressing, linear probing.  keys: EMPTY / TOMB / id.  vals: FREE (all ones) when
// the key is not committed, PENDING|batch_index while an enqueue batch is being resolved,
// (pool_generation << 32 | pool_slot) once the player is queued.
// =======================================================================================

// E1: validate + claim.  The lowest batch index wins a repeated id (atomicMin), which
// is what a serialized in_queue?/add_user sequence (middleware/worker.ex:65-70) yields.
__global__ void k_enq_claim(uint32_t base, uint32_t n, const uint64_t* __restrict__ id, const int32_t* __restrict__ rating,
                            const uint8_t* __restrict__ mode, const uint8_t* __restrict__ grp_lut, int32_t key_lo,
                            uint32_t KR, uint32_t n_modes, ActiveView act, uint64_t* __restrict__ hslot,
                            uint8_t* __restrict__ code) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // this launch covers batch indices [base, base + n)
  if (t >= n) return;
  const uint32_t i = base + t;
  const uint64_t pid = id[i];
  const int32_t hi = key_lo + (int32_t)KR - 1;
  const int32_t r = rating[i] < key_lo ? key_lo : (rating[i] > hi ? hi : rating[i]);
  if (mode[i] >= n_modes || pid >= kTombKey || grp_lut[r - key_lo] == 0xFF) { code[i] = 2; hslot[i] = ~0ull; return; }
  if (!act.mask) { code[i] = 1; hslot[i] = ~0ull; return; }
  uint64_t h = hash64(pid) & act.mask;
  for (uint64_t probe = 0; probe <= act.mask; ++probe) {
    unsigned long long k = act.keys[h];
    if (k == kEmptyKey) {
      k = atomicCAS(&act.keys[h], kEmptyKey, pid);
      if (k == kEmptyKey) k = pid;
    }
    if (k == pid) {
      const unsigned long long old = atomicMin(&act.vals[h], kPending | i);
      code[i] = (old < kPending) ? 0 : 1;  // committed entry -> "already in the queue"
      hslot[i] = h;
      return;
    }
    h = (h + 1) & act.mask;
  }
  code[i] = 3; hslot[i] = ~0ull;  // table full
}

// E2: winners = entries whose PENDING index is their own; per-block winner counts.
// E2 / E3 run per ingest chunk — batch indices [base, base + n) — so that they overlap the
// host-to-device copy of the next chunk; the lowest batch index wins a repeated id, and a chunk's
// winners are final once every lower index has claimed.
__global__ void k_enq_count(uint32_t base, uint32_t n, ActiveView act, const uint64_t* __restrict__ hslot,
                            uint8_t* __restrict__ code, uint32_t* __restrict__ blocksum) {
  __shared__ uint32_t s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = base + t;
  bool win = false;
  if (t < n && code[i] == 1) {
    win = !act.mask || act.vals[hslot[i]] == (kPending | i);
    if (!win) code[i] = 0;  // a lower batch index holds the id
  }
  const uint32_t b = __ballot_sync(0xFFFFFFFFu, win);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_cnt, (uint32_t)__popc(b));
  __syncthreads();
  if (threadIdx.x == 0) blocksum[blockIdx.x] = s_cnt;
}

// exclusive scan of blocksum (single CTA; nblocks is at most a few 10k), continued from the
// running total of the earlier chunks of the batch (*total), which it then advances
__global__ void __launch_bounds__(1024) k_scan_small(uint32_t nb, uint32_t* __restrict__ v, uint32_t* __restrict__ total) {
  __shared__ uint32_t s_sum[1024];
  const uint32_t tid = threadIdx.x;
  const uint32_t before = *total;
  const uint32_t per = (nb + 1023) / 1024;
  const uint32_t lo = tid * per, hi = (lo + per < nb) ? lo + per : nb;
  uint32_t local = 0;
  for (uint32_t i = lo; i < hi && i < nb; ++i) local += v[i];
  s_sum[tid] = local;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    uint32_t x = (tid >= (uint32_t)off) ? s_sum[tid - off] : 0;
    __syncthreads();
    s_sum[tid] += x;
    __syncthreads();
  }
  uint32_t run = before + s_sum[tid] - local;
  for (uint32_t i = lo; i < hi && i < nb; ++i) { const uint32_t x = v[i]; v[i] = run; run += x; }
  __syncthreads();  // everyone has read *total
  if (tid == 1023) *total = before + s_sum[1023];
}

// E3: append winners to the pool in batch order (= enqueue order) and commit their
// active-set entries.  Players past the pool capacity are rolled back with code 3.
__global__ void k_enq_append(uint32_t base, uint32_t n, const uint64_t* __restrict__ id, const int32_t* __restrict__ rating,
                             const uint8_t* __restrict__ mode, const uint32_t* __restrict__ ts,
                             const uint8_t* __restrict__ mode_tsize, ActiveView act, const uint64_t* __restrict__ hslot,
                             uint8_t* __restrict__ code, const uint32_t* __restrict__ blockoff, PoolView pool,
                             uint32_t n_pool, uint32_t capacity, uint32_t gen, uint32_t* __restrict__ n_rejected_cap,
                             BinMap bm) {
  __shared__ uint32_t s_warp[32];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = base + t;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool win = t < n && code[i] == 1;
  const uint32_t b = __ballot_sync(0xFFFFFFFFu, win);
  if (lane == 0) s_warp[warp] = __popc(b);
  __syncthreads();
  uint32_t woff = 0;
  for (uint32_t w = 0; w < warp; ++w) woff += s_warp[w];
  if (!win) return;
  const uint32_t slot = n_pool + blockoff[blockIdx.x] + woff + __popc(b & ((1u << lane) - 1u));
  if (slot >= capacity) {
    code[i] = 3;
    if (act.mask) { act.vals[hslot[i]] = kFreeVal; act.keys[hslot[i]] = kTombKey; }
    atomicAdd(n_rejected_cap, 1u);
    return;
  }
  pool.id[slot] = id[i]; pool.rating[slot] = rating[i]; pool.mode[slot] = mode[i];
  pool.tsize[slot] = mode_tsize[mode[i]]; pool.ts[slot] = ts ? ts[i] : 0u;
  pool.bin[slot] = (uint16_t)bin_of(bm, bm.lut, rating[i], mode[i]);  // the tick's sort key, derived once at ingest
  if (act.mask) act.vals[hslot[i]] = ((unsigned long long)gen << 32) | slot;
}

// ActiveUser.remove_user/1 (models/active_user.ex:57-66), batched.  A player still
// queued is tombstoned in the pool (mode byte = DEAD) so the next tick drops it the way
// remove_inactive_players/1 (search/worker.ex:267-280) filters it.
__global__ void k_remove(uint32_t n, const uint64_t* __restrict__ id, ActiveView act, PoolView pool, uint32_t n_pool,
                         uint32_t gen, uint32_t dead_bin, uint32_t* __restrict__ n_removed) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !act.mask) return;
  const uint64_t pid = id[i];
  if (pid >= kTombKey) return;
  uint64_t h = hash64(pid) & act.mask;
  for (uint64_t probe = 0; probe <= act.mask; ++probe) {
    const unsigned long long k = act.keys[h];
    if (k == kEmptyKey) return;
    if (k == pid) {
      const unsigned long long v = act.vals[h];
      if (atomicCAS(&act.keys[h], (unsigned long long)pid, kTombKey) != pid) return;  // a twin in this batch won
      act.vals[h] = kFreeVal;
      const uint32_t slot = (uint32_t)v, g = (uint32_t)(v >> 32);
      if (v < kPending && g == gen && slot < n_pool && pool.id[slot] == pid) {
        pool.mode[slot] = MM_MODE_DEAD;
        pool.bin[slot] = (uint16_t)dead_bin;
      }
      atomicAdd(n_removed, 1u);
      return;
    }
    h = (h + 1) & act.mask;
  }
}

// ActiveUser.in_queue?/1 (models/active_user.ex:33-44), batched.
__global__ void k_lookup(uint32_t n, const uint64_t* __restrict__ id, ActiveView act, uint8_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t pid = id[i];
  uint8_t found = 0;
  if (act.mask && pid < kTombKey) {
    uint64_t h = hash64(pid) & act.mask;
    for (uint64_t probe = 0; probe <= act.mask; ++probe) {
      const unsigned long long k = act.keys[h];
      if (k == kEmptyKey) break;
      if (k == pid) { found = 1; break; }
      h = (h + 1) & act.mask;
    }
  }
  out[i] = found;
}

// Rebuild without tombstones: re-insert every committed entry of the old table.
__global__ void k_rehash(ActiveView oldt, ActiveView newt) {
  for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s <= oldt.mask; s += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = oldt.keys[s];
    if (k >= kTombKey) continue;
    uint64_t h = hash64(k) & newt.mask;
    for (;;) {
      if (atomicCAS(&newt.keys[h], kEmptyKey, k) == kEmptyKey) { newt.vals[h] = oldt.vals[s]; break; }
      h = (h + 1) & newt.mask;
    }
  }
}

// After mm_restore: point every queued player's entry at its slot again.
__global__ void k_restamp(PoolView pool, uint32_t n_pool, ActiveView act, uint32_t gen) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pool || !act.mask) return;
  if (pool.mode[i] == MM_MODE_DEAD) return;
  const uint64_t pid = pool.id[i];
  uint64_t h = hash64(pid) & act.mask;
  for (uint64_t probe = 0; probe <= act.mask; ++probe) {
    const unsigned long long k = act.keys[h];
    if (k == kEmptyKey) return;
    if (k == pid) { act.vals[h] = ((unsigned long long)gen << 32) | i; return; }
    h = (h + 1) & act.mask;
  }
}

// empty active set: every slot {EMPTY key, FREE value}
__global__ void k_fill_kv(ulonglong2* p, uint64_t n, unsigned long long k, unsigned long long v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    p[i] = make_ulonglong2(k, v);
}

}  // namespace mm
