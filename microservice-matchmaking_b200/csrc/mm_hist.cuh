// mm_hist.cuh — phase 1 of the tick: row histograms of the resident bin column
#pragma once
#include "mm_common.cuh"

namespace mm {

// ---------------------------------------------------------------------------------------
// hist_body<BLOCK>: M[row][bin] = number of the row's players in that bin, straight from the resident 16-bit
// bin column (the sort key, derived once at ingest by k_enq_append, tombstoned by k_remove, carried through
// the compaction).  The row's tiles stream through a ring of TMA bulk copies (one 4 KB chunk of the bin
// column per tile, L2 evict-last: the placement pass re-reads the column from L2); shared-memory atomics
// build the row histogram.  Only the bins of the partitions the row touches are written to M — a row's tiles
// are consecutive in (partition, chunk) order, so that is one contiguous bin range.
// Shared memory: ring[8][kTile] u16 | mbarriers | nvalid[8] | hist[Kp] | tile descriptors.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kHistStages = 8;  // ring depth: two half-turns of 4 tiles (8 192 players each)

template <int BLOCK>
__device__ __forceinline__ void hist_body(unsigned char* smem_raw, const Geo& g, const uint16_t* __restrict__ bins16,
                                          const PoolMeta meta, uint32_t Kp,
                                          const uint32_t* __restrict__ seg_bin_lo, uint32_t* __restrict__ M) {
  constexpr uint32_t kBytes = kTile * 2, S = kHistStages, H = S / 2;
  uint16_t* ring = reinterpret_cast<uint16_t*>(smem_raw);                            // [S][kTile]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S * kBytes);       // [S]
  uint32_t* s_nv = reinterpret_cast<uint32_t*>(smem_raw + (size_t)S * kBytes + 64);  // [S]
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + (size_t)S * kBytes + 128); // [Kp]
  DescCache& dc = *reinterpret_cast<DescCache*>(hist + ((Kp + 3) & ~3u));
  const uint32_t tid = threadIdx.x, row = blockIdx.x;
  const uint64_t pol = policy_evict_last();
  const uint32_t s0 = row * g.tpr < g.NT ? row * g.tpr : g.NT;
  const uint32_t s1 = s0 + g.tpr < g.NT ? s0 + g.tpr : g.NT;
  const uint32_t n_tiles = s1 - s0;
  if (tid == 0) {
    for (uint32_t s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  fence_proxy_async();
  desc_fill<BLOCK>(dc, g, meta, s0, s1);
  __syncthreads();
  uint32_t dbase = 0;  // first row tile covered by the descriptor cache
  auto issue = [&](uint32_t stage, uint32_t t) {  // thread 0: the tile's bulk copy
    uint32_t phys, nv;
    if (t - dbase < kDescCap) { phys = dc.phys[t - dbase]; nv = dc.nvsg[t - dbase] & 0xFFFFu; }
    else { const TileDesc d = geo_tile(g, meta, s0 + t); phys = d.phys; nv = d.nvalid; }  // ahead of the cache
    s_nv[stage] = nv;
    mbar_expect_tx(&full[stage], kBytes);
    tma_load_1d(ring + (size_t)stage * kTile, bins16 + (size_t)phys * kTile, kBytes, &full[stage], pol);
  };
  if (tid == 0)
    for (uint32_t t = 0; t < S && t < n_tiles; ++t) issue(t, t);
  for (uint32_t i = tid; i < Kp; i += BLOCK) hist[i] = 0;
  __syncthreads();
  // The ring is consumed half a turn at a time (4 tiles = up to 8 192 players between two CTA barriers) while the
  // other half's copies are in flight: the pass is latency-bound at this size, fewer and fatter steps win.
  for (uint32_t i = 0; i * H < n_tiles; ++i) {
    const uint32_t t0 = i * H, sb = (i & 1u) * H, parity = (i >> 1) & 1u;
    if (t0 >= dbase + kDescCap) {  // (uniform) next batch of descriptors; thread 0 is not issuing right now
      dbase = t0;
      desc_fill<BLOCK>(dc, g, meta, s0 + t0, s1);
      __syncthreads();
    }
    const uint32_t nt = n_tiles - t0 < H ? n_tiles - t0 : H;
    for (uint32_t k = 0; k < nt; ++k) {
      const uint16_t* tb = ring + (size_t)(sb + k) * kTile;
      mbar_wait(&full[sb + k], parity);
      const uint32_t valid = s_nv[sb + k];
#pragma unroll
      for (uint32_t q = tid; q < kTile / 8; q += BLOCK) {  // 8 bins (128 bits) per thread per step
        const uint32_t o = q * 8;
        if (o + 8 <= valid) {
          const uint4 v = *reinterpret_cast<const uint4*>(tb + o);
          atomicAdd(&hist[v.x & 0xFFFFu], 1u); atomicAdd(&hist[v.x >> 16], 1u);
          atomicAdd(&hist[v.y & 0xFFFFu], 1u); atomicAdd(&hist[v.y >> 16], 1u);
          atomicAdd(&hist[v.z & 0xFFFFu], 1u); atomicAdd(&hist[v.z >> 16], 1u);
          atomicAdd(&hist[v.w & 0xFFFFu], 1u); atomicAdd(&hist[v.w >> 16], 1u);
        } else {
          for (uint32_t e = o; e < valid; ++e) atomicAdd(&hist[tb[e]], 1u);
        }
      }
    }
    __syncthreads();
    if (tid == 0)
      for (uint32_t k = 0; k < H && t0 + S + k < n_tiles; ++k) issue(sb + k, t0 + S + k);
  }
  if (n_tiles) {
    const uint32_t p_first = geo_seg_of(g, s0), p_last = geo_seg_of(g, s1 - 1);
    const uint32_t blo = seg_bin_lo[p_first], bhi = seg_bin_lo[p_last + 1];
    uint32_t* mrow = M + (size_t)row * Kp;
    for (uint32_t i = blo + tid; i < bhi; i += BLOCK) mrow[i] = hist[i];
  }
  if (tid == 0)
    for (uint32_t s = 0; s < kHistStages; ++s) mbar_inval(&full[s]);
}

__host__ __device__ constexpr size_t hist_smem_bytes(uint32_t Kp) {
  return (size_t)kHistStages * kTile * 2 + 128 + (size_t)((Kp + 3) & ~3u) * 4 + sizeof(DescCache) + 16;
}

}  // namespace mm
