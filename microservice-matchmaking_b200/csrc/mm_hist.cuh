// mm_hist.cuh — phase 1 of the tick: row histograms (k_hist3; k_hist = first version, cross-check)
#pragma once
#include "mm_common.cuh"

namespace mm {

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

}  // namespace mm
