// mm_place.cuh — phase 3 of the tick: stable rank inside the row and the id scatter (k_place2; k_place<0|1> = first versions, cross-checks)
#pragma once
#include "mm_common.cuh"
#include "mm_scan.cuh"

namespace mm {

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

}  // namespace mm
