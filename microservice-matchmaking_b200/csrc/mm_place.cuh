// mm_place.cuh — phase 3 of the tick: stable rank inside the row -> final lobby-major slot of every player
#pragma once
#include "mm_common.cuh"

namespace mm {

// ---------------------------------------------------------------------------------------
// place_body<BLOCK>: the placement pass.  The row (= CTA) streams its tiles — 2 048 players of ONE (mode, group)
// partition each: (bin u16, id u64) chunks through a ring of TMA bulk copies (cp.async.bulk -> mbarrier),
// issued `stages` tiles ahead by one thread, L2 evict-first — computes every player's STABLE rank among the row's
// players of the same bin and stores the id to its final lobby-major slot
//     slot = outbase[bin] + (players of the bin in earlier rows: M) + rank inside the row.
// A player at or past binlim[bin] stays queued: one bit in left_bits (one ballot per 32 players, plain word stores).
//
// Two ranking paths, chosen per tile (uniform for the CTA):
//  * FAST — the tile's partition has <= 255 bins (e.g. 5 001 rating values in 32 groups: 157): a one-pass 8-bit
//    counting sort of the tile in shared memory.  Warp w owns 128 consecutive tile positions; per 32 players the
//    peers with the same digit are found with <= 8 ballots (MATCH.ANY costs 64 cycles per warp instruction on
//    B200), the lowest peer bumps the warp's private digit counter; 128 threads then scan the 16 x 256 counter
//    matrix (two 16-bit digits per word, packed adds) into tile-local sorted positions.  Ids are staged AT THEIR
//    SORTED POSITION in the tile's own ring stage together with their global slot, and the CTA writes the staged
//    tile back in sorted order: consecutive threads store consecutive slots of a bin's run (about 2048 / bins ids =
//    100+ contiguous bytes), so member_ids is written in whole 32-byte sectors instead of 8-byte fragments.
//  * LIST — larger key domains (one group of 5 001 rating values): one list node per player on a hashed,
//    epoch-tagged head table (shared-memory atomicExch), walk of the round-local list counting same-bin nodes with a
//    smaller tile position; ids are scattered straight from registers.  `heavy` ticks (some bin expects > 8
//    players per tile) aggregate the nodes per warp first (__match_any_sync) so a list never exceeds 64 nodes.
//
// Shared memory: ring_ids[stages][kTile] | ring_bins[stages][kTile] | mbarriers + tile descriptors | cnt[keys of one partition] |
//                union { LIST: head[kHeadSlots] node[kTile] nbin[kTile] ; FAST: wcnt[16][256] u16, sslot[kTile] }.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kHeadSlots = 4096;
constexpr uint32_t NW16 = 16;  // warps per 512-thread CTA
constexpr uint32_t kPlaceUnionBytes = NW16 * 256 * 4 + NW16 * 256 * 2 + 1024 + kTile * 4;  // FAST: 16 + 8 + 1 + 8 KB >= LIST: 28 KB

// max_nb = most sort keys any one partition has: the slot counters are kept per partition, not for the whole key domain
__host__ __device__ constexpr uint32_t place_cnt_cap(uint32_t max_nb) { return (max_nb > 1024u ? max_nb : 1024u); }
__host__ __device__ constexpr size_t place_smem_bytes(uint32_t max_nb, uint32_t stages) {
  return (size_t)stages * kTileBytes + 128 + (size_t)((place_cnt_cap(max_nb) + 3) & ~3u) * 4 + kPlaceUnionBytes + sizeof(DescCache) + 16;
}

struct PlaceArgs {
  const uint16_t* bins16;
  const uint64_t* ids;
  PoolMeta meta;
  uint32_t K, Kp, R, stages, fast_ok, max_nb;
  const uint32_t* seg_bin_lo;
  const uint16_t* bin_seg;
  const uint32_t* M;     // [rows][Kp] row histograms (raw counts)
  const uint32_t* P;     // [rows][Kp] their column prefixes, only when the column-scan phase ran
  const uint32_t* outbase;
  const uint32_t* binlim;
  uint64_t* members;
  uint32_t* src_idx;    // optional: virtual pool position of every member (emission order in ARRIVAL mode)
  uint32_t* left_bits;  // one bit per virtual pool position: the player stays queued after this tick
  uint32_t* rescnt;     // [R] players of the row that stay queued
  TickCtr* ctr;
};

template <int BLOCK>
__device__ __forceinline__ void place_body(unsigned char* smem_raw, const Geo& g, const PlaceArgs a) {
  static_assert(BLOCK == 512 && kTile == 2048, "tile arrangement is written for 512 threads x 4 players");
  constexpr int J = kTile / BLOCK;
  constexpr int NW = BLOCK / 32;
  const uint32_t stages = a.stages, K = a.K, Kp = a.Kp;
  uint64_t* ring_ids = reinterpret_cast<uint64_t*>(smem_raw);                               // [stages][kTile]
  uint16_t* ring_bins = reinterpret_cast<uint16_t*>(smem_raw + (size_t)stages * kTile * 8);  // [stages][kTile]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * kTileBytes);      // [kMaxStages]
  uint32_t* s_nv = reinterpret_cast<uint32_t*>(smem_raw + (size_t)stages * kTileBytes + 32); // [kMaxStages] valid players
  uint32_t* s_sg = s_nv + kMaxStages;                                                        // [kMaxStages] partition
  uint32_t* s_b0 = s_sg + kMaxStages;                                                        // [kMaxStages] its first bin
  uint32_t* s_b1 = s_b0 + kMaxStages;                                                        // [kMaxStages] its end bin
  uint32_t* s_misc = s_b1 + kMaxStages;                                                      // [8]
  uint32_t* cnt = reinterpret_cast<uint32_t*>(smem_raw + (size_t)stages * kTileBytes + 128); // [max_nb] of the current partition
  const uint32_t cnt_cap = place_cnt_cap(a.max_nb);
  unsigned char* uni = reinterpret_cast<unsigned char*>(cnt + ((cnt_cap + 3) & ~3u));
  DescCache& dc = *reinterpret_cast<DescCache*>(uni + kPlaceUnionBytes);
  // LIST
  uint32_t* head = reinterpret_cast<uint32_t*>(uni);           // [kHeadSlots]
  uint32_t* node = head + kHeadSlots;                          // [kTile]
  uint16_t* nbin = reinterpret_cast<uint16_t*>(node + kTile);  // [kTile] heavy path: bin of a group node
  // FAST
  uint32_t* wmask = reinterpret_cast<uint32_t*>(uni);          // [NW][256] per-warp match masks (all-zero between items)
  uint16_t* wcnt = reinterpret_cast<uint16_t*>(wmask + NW * 256);  // [NW][256] per-warp running digit counters
  uint32_t* wcnt32 = reinterpret_cast<uint32_t*>(wcnt);        // the same, two digits per word: [NW][128]
  uint32_t* lgd = wcnt32 + NW * 128;                           // [256] (global slot base - tile-local base) | flag
  uint32_t* spd = lgd + 256;                                   // [kTile] sorted position -> tile position | digit << 11

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t row = blockIdx.x;
  const uint64_t pol_in = policy_evict_first();

  const uint32_t s0 = row * g.tpr < g.NT ? row * g.tpr : g.NT;
  const uint32_t s1 = s0 + g.tpr < g.NT ? s0 + g.tpr : g.NT;
  const uint32_t n_tiles = s1 - s0;

  uint32_t dbase = 0;  // first row tile covered by the descriptor cache
  auto issue = [&](uint32_t stage, uint32_t t) {  // thread 0: descriptor + the tile's two bulk copies
    uint32_t phys, nvsg;
    if (t - dbase < kDescCap) { phys = dc.phys[t - dbase]; nvsg = dc.nvsg[t - dbase]; }
    else { const TileDesc d = geo_tile(g, a.meta, s0 + t); phys = d.phys; nvsg = d.nvalid | (d.seg << 16); }
    s_nv[stage] = nvsg & 0xFFFFu;
    s_sg[stage] = nvsg >> 16;
    s_b0[stage] = a.seg_bin_lo[nvsg >> 16];
    s_b1[stage] = a.seg_bin_lo[(nvsg >> 16) + 1];
    mbar_expect_tx(&full[stage], kTileBytes);
    tma_load_1d(ring_ids + (size_t)stage * kTile, a.ids + (size_t)phys * kTile, kTile * 8, &full[stage], pol_in);
    tma_load_1d(ring_bins + (size_t)stage * kTile, a.bins16 + (size_t)phys * kTile, kTile * 2, &full[stage], pol_in);
  };

  if (tid == 0) {
    for (uint32_t s = 0; s < stages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
    s_misc[0] = 0;  // players of the row that stay queued
    s_misc[2] = 0;  // FAST: the current tile has players past their bin's matched prefix
  }
  fence_proxy_async();
  desc_fill<BLOCK>(dc, g, a.meta, s0, s1);
  __syncthreads();
  if (tid == 0)
    for (uint32_t t = 0; t < stages && t < n_tiles; ++t) issue(t, t);
  for (uint32_t i = tid; i < kHeadSlots + NW * 128; i += BLOCK) head[i] = 0;  // LIST heads = FAST mask table; + FAST counters
  const bool heavy = __ldcg(&a.ctr->heavy) != 0;
  __syncthreads();

  uint32_t st = 0, parity = 0;
  uint32_t nleft = 0;  // lane 0: players of this warp's positions that stay queued
  uint32_t wb = 0, we = 0;  // [wb, we): bins whose slot counters are loaded
  const uint32_t row_p_last = n_tiles ? geo_seg_of(g, s1 - 1) : 0u;
  uint32_t uni_st = 0;  // (uniform) who dirtied the union region: 0 nobody (all-zero), 1 LIST, 2 FAST
  for (uint32_t t = 0; t < n_tiles; ++t) {
    if (t == dbase + kDescCap) {  // (uniform) next batch of descriptors; thread 0 is not issuing right now
      dbase = t;
      desc_fill<BLOCK>(dc, g, a.meta, s0 + t, s1);
      __syncthreads();
    }
    const uint32_t vbase = (s0 + t) * kTile;  // virtual position of the tile's first player
    uint16_t* tb = ring_bins + (size_t)st * kTile;
    uint64_t* ti = ring_ids + (size_t)st * kTile;
    mbar_wait(&full[st], parity);
    const uint32_t valid = s_nv[st];
    const uint32_t bin0 = s_b0[st], nb = s_b1[st] - bin0;
    if (bin0 < wb || bin0 + nb > we) {
      // (uniform) the row enters a partition whose slot counters are not loaded: load a window of whole partitions
      // starting with this one (a row's tiles come in partition order; a row usually spans 1-3 partitions, which fit
      // at once).  cnt[b - wb] = slot of the (row, b) cell's first player; bit 31 flags a cell that reaches past
      // the bin's matched prefix (only those players look at binlim).
      wb = bin0; we = bin0 + nb;
      for (uint32_t p = s_sg[st] + 1; p <= row_p_last; ++p) {
        const uint32_t e = a.seg_bin_lo[p + 1];
        if (e - wb > cnt_cap) break;
        we = e;
      }
      const bool scanned = geo_use_colscan(g);  // the column-scan phase ran: P holds the row prefixes
      __syncthreads();
      for (uint32_t i = wb + tid; i < we; i += BLOCK) {
        uint32_t rlo = 0, rhi = 0, v = 0;
        if (geo_rows_of(g, a.bin_seg[i], rlo, rhi) && row >= rlo && row <= rhi) {
          // __ldcg: these arrays are produced earlier in the same (fused) launch by other SMs
          uint32_t pre = 0;
          if (scanned) pre = __ldcg(&a.P[(size_t)row * Kp + i]);
          else
            for (uint32_t r = rlo; r < row; r += 8) {  // few rows per partition; 8 independent L2 loads in flight
              uint32_t v[8];
#pragma unroll
              for (uint32_t u = 0; u < 8; ++u) v[u] = r + u < row ? __ldcg(&a.M[(size_t)(r + u) * Kp + i]) : 0u;
#pragma unroll
              for (uint32_t u = 0; u < 8; ++u) pre += v[u];
            }
          const uint32_t c = __ldcg(&a.M[(size_t)row * Kp + i]);
          const uint32_t start = __ldcg(&a.outbase[i]) + pre;
          v = start | ((start + c > __ldcg(&a.binlim[i])) ? 0x80000000u : 0u);
        }
        cnt[i - wb] = v;
      }
      __syncthreads();
    }
    uint32_t* cntp = cnt + (bin0 - wb);  // counters of this tile's partition
    const bool fast = a.fast_ok && nb <= kFastBins;
    uint32_t lmask = 0;  // bit j: my j-th player stays queued

    if (fast) {
      // ---------------- FAST: 8-bit counting sort of the tile in shared memory ----------------
      if (uni_st == 1) {  // the LIST path left head[] entries in the mask / counter tables
        for (uint32_t i = tid; i < NW * 256; i += BLOCK) wmask[i] = 0;
        for (uint32_t i = tid; i < NW * 128; i += BLOCK) wcnt32[i] = 0;
      }
      uint32_t dg[J], rk[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {  // warp-striped: position = warp * 128 + j * 32 + lane
        const uint32_t pos = warp * (32 * J) + j * 32 + lane;
        const uint32_t d = (uint32_t)tb[pos] - bin0;  // digits 0 .. nb-1 live, nb = dead / past the tile's end
        dg[j] = (pos < valid && d < nb) ? d : nb;
      }
      if (uni_st == 1) __syncthreads();  // (uniform) the tables were just re-zeroed
      uni_st = 2;
      {   // the counter table is all-zero here (previous tile / prologue), the mask table always is between items
        // Peers of the same digit among the warp's 32 players: every lane ORs its bit into the warp's mask table
        // (shared-memory RED), reads the word back — that IS the match mask — and the peers reset the word
        // and bump the warp's running digit counter.  3 shared-memory instructions per 32 players instead of 8
        // ballots + selects (MATCH.ANY costs 64 cycles per warp instruction on B200).
        uint32_t* wm = wmask + warp * 256;
        uint16_t* wc = wcnt + warp * 256;
        const uint32_t lbit = 1u << lane;
        if (nb > 16) {
#pragma unroll
          for (int j = 0; j < J; ++j) {
            atomicOr(&wm[dg[j]], lbit);
            __syncwarp();
            const uint32_t peers = wm[dg[j]];
            const uint32_t base = wc[dg[j]];
            __syncwarp();
            wm[dg[j]] = 0;                                    // every peer stores the same values: no leader election,
            wc[dg[j]] = (uint16_t)(base + __popc(peers));     // no divergence
            __syncwarp();
            rk[j] = base + __popc(peers & lt_mask);
          }
        } else {
          // a handful of keys (arrival order: ONE per partition): the 32 lanes would serialise on a few mask words,
          // so the peers come from <= 5 ballots instead
          const uint32_t nbits = 32u - __clz(nb);
#pragma unroll
          for (int j = 0; j < J; ++j) {
            uint32_t peers = 0xFFFFFFFFu;
#pragma unroll
            for (uint32_t bit = 0; bit < 5; ++bit)
              if (bit < nbits) {
                const bool on = (dg[j] >> bit) & 1u;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, on);
                peers &= on ? bal : ~bal;
              }
            const uint32_t base = wc[dg[j]];
            __syncwarp();
            wc[dg[j]] = (uint16_t)(base + __popc(peers));
            __syncwarp();
            rk[j] = base + __popc(peers & lt_mask);
          }
        }
      }
      __syncthreads();  // B2: per-warp digit counts complete
      if (tid < 128) {
        // digits 2*tid, 2*tid+1: column scan over the 16 warps (packed 16-bit adds), exclusive scan over the digits,
        // global slot base from the row's running bin counters
        uint32_t v[NW], sum = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) { v[w] = wcnt32[w * 128 + tid]; sum += v[w]; }
        const uint32_t lo = sum & 0xFFFFu, hi = sum >> 16, both = lo + hi;
        uint32_t incl = both;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, off);
          if (lane >= (uint32_t)off) incl += u;
        }
        if (lane == 31) s_misc[4 + warp] = incl;
        bar_sync_named(128);
        uint32_t wbase = 0;
        for (uint32_t w = 0; w < warp; ++w) wbase += s_misc[4 + w];
        const uint32_t l0 = wbase + incl - both, l1 = l0 + lo;  // tile-local sorted position of the digits' first players
        const uint32_t d0 = 2 * tid, d1 = d0 + 1;
        uint32_t fl = 0;
        if (d0 < nb) { const uint32_t base = cntp[d0]; cntp[d0] = base + lo; lgd[d0] = (((base & 0x7FFFFFFFu) - l0) & 0x7FFFFFFFu) | (base & 0x80000000u); if (lo) fl |= base; }
        if (d1 < nb) { const uint32_t base = cntp[d1]; cntp[d1] = base + hi; lgd[d1] = (((base & 0x7FFFFFFFu) - l1) & 0x7FFFFFFFu) | (base & 0x80000000u); if (hi) fl |= base; }
        if (fl >> 31) s_misc[2] = 1;   // some player of this tile sits in a cell that reaches past its bin's matched prefix
        if (d0 == nb) s_misc[1] = l0;  // live players of the tile (the dead digit sorts last)
        if (d1 == nb) s_misc[1] = l1;
        uint32_t run = l0 | (l1 << 16);
#pragma unroll
        for (int w = 0; w < NW; ++w) { wcnt32[w * 128 + tid] = run; run += v[w]; }
      }
      __syncthreads();  // B3: wcnt[w][d] = tile-local sorted position of warp w's first player of digit d
      const bool anyf = s_misc[2] != 0;  // (uniform)
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (dg[j] < nb) {
          const uint32_t lpos = wcnt[warp * 256 + dg[j]] + rk[j];
          bool matched = true;
          if (anyf || a.src_idx) {
            const uint32_t e = lgd[dg[j]];
            const uint32_t slot = (e + lpos) & 0x7FFFFFFFu;
            if (e >> 31) matched = slot < __ldcg(&a.binlim[bin0 + dg[j]]);
            if (matched && a.src_idx) a.src_idx[slot] = vbase + warp * (32 * J) + j * 32 + lane;
          }
          // sorted position -> (tile position, digit); 255 = stays queued.  The ids stay where the TMA put them.
          spd[lpos] = (warp * (32 * J) + j * 32 + lane) | ((matched ? dg[j] : 255u) << 11);
          if (!matched) lmask |= 1u << j;
        }
      }
      {  // left_bits: the warp owns 128 consecutive positions = 4 words; lane j stores word j
        uint32_t mine = 0, all = 0;
        if (anyf) {
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const uint32_t wv = __ballot_sync(0xFFFFFFFFu, (lmask >> j) & 1u);
            if (lane == (uint32_t)j) mine = wv;
            all += __popc(wv);
          }
        }
        if (lane < (uint32_t)J) a.left_bits[(vbase >> 5) + warp * J + lane] = mine;
        if (lane == 0) nleft += all;
      }
      __syncthreads();  // B4: the sorted order of the tile is staged
      {
        const uint32_t n_live = s_misc[1];
#pragma unroll
        for (int i = 0; i < J; ++i) {
          const uint32_t k = i * BLOCK + tid;
          if (k < n_live) {
            const uint32_t v = spd[k], d = v >> 11;
            if (d != 255u) a.members[(lgd[d] + k) & 0x7FFFFFFFu] = ti[v & 0x7FFu];
          }
        }
        reinterpret_cast<uint4*>(wcnt32)[tid] = make_uint4(0, 0, 0, 0);  // counter table all-zero again (8 KB = 512 x 16 B)
        if (tid == 0) s_misc[2] = 0;
      }
    } else {
      // ---------------- LIST: hashed per-bin lists, ids scattered from registers ----------------
      if (uni_st == 2) {  // the FAST path left counters / slot bases in the head table
        for (uint32_t i = tid; i < kHeadSlots; i += BLOCK) head[i] = 0;
        __syncthreads();
      }
      uni_st = 1;
      const uint32_t epoch = t + 1;
      uint32_t bin[J], slot[J];
      uint64_t idv[J];
      bool flag[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {  // strided: position = j * BLOCK + tid
        const uint32_t pos = j * BLOCK + tid;
        bin[j] = (pos < valid) ? (uint32_t)tb[pos] : 0xFFFFu;
        slot[j] = 0; flag[j] = false;
      }
      if (!heavy) {
        uint32_t snap[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const uint32_t pos = j * BLOCK + tid;
          snap[j] = 0;
          if (bin[j] < K) {
            snap[j] = cntp[bin[j] - bin0];
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
            if (lower == 0) cntp[bin[j] - bin0] = snap[j] + total;  // the bin's earliest player of the tile
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
            snap[j] = cntp[bin[j] - bin0];
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
            if (lower == 0) cntp[bin[j] - bin0] = snap[j] + total;
          }
          bg = __shfl_sync(0xFFFFFFFFu, bg, leader[j]);
          slot[j] = (bg & 0x7FFFFFFFu) + rankw[j];
          flag[j] = (bg >> 31) != 0;
        }
#pragma unroll
        for (int j = 0; j < J; ++j) idv[j] = ti[j * BLOCK + tid];
      }
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (bin[j] < K) {
          bool matched = true;
          if (flag[j]) matched = slot[j] < __ldcg(&a.binlim[bin[j]]);
          if (matched) {
            a.members[slot[j]] = idv[j];
            if (a.src_idx) a.src_idx[slot[j]] = vbase + j * BLOCK + tid;
          } else {
            lmask |= 1u << j;
          }
        }
      }
      {  // strided: batch j of the warp = positions j*BLOCK + 32*warp .. +31 = one word; lane j stores batch j's word
        uint32_t mine = 0, all = 0;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const uint32_t wv = __ballot_sync(0xFFFFFFFFu, (lmask >> j) & 1u);
          if (lane == (uint32_t)j) mine = wv;
          all += __popc(wv);
        }
        if (lane < (uint32_t)J) a.left_bits[(vbase + lane * BLOCK + warp * 32) >> 5] = mine;
        if (lane == 0) nleft += all;
      }
    }
    __syncthreads();  // everyone is done with stage st and with this tile's ranking state
    if (tid == 0 && t + stages < n_tiles) issue(st, t + stages);
    if (++st == stages) { st = 0; parity ^= 1u; }
  }

  if (lane == 0 && nleft) atomicAdd(&s_misc[0], nleft);
  __syncthreads();
  if (tid == 0) {
    a.rescnt[row] = s_misc[0];  // players of this row that stay queued
    for (uint32_t s = 0; s < stages; ++s) mbar_inval(&full[s]);
  }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 2) k_place(const PlaceArgs a, const uint32_t* __restrict__ fill, uint32_t n_segs) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ Geo geo;
  __shared__ uint32_t s_gtmp[33];
  geo_build<BLOCK>(geo, fill, n_segs, a.R, s_gtmp);
  place_body<BLOCK>(smem_raw, geo, a);
}

}  // namespace mm
