// mm_kernels.cuh — device code of the search tick (sm_100a).
//
// The tick replaces, for every queued player at once, the per-request loop of
// Search.Worker.consume/5 (reference matchmaking/lib/search/worker.ex:291-324).
// Under the serialized oracle (oracle/mm_oracle.c) that loop has the closed form
//   "drop inactive players, stable-partition the feed order by (mode, group), cut
//    each partition into lobbies of L"
// which on the GPU is ONE stable counting sort over a small key domain:
//   bin(player) = mode * stride + lut[clamp(rating)]          (K bins, K ~ 5k * modes)
// followed by a per-(mode, group)-segment cut.  Kernels:
//   k_hist     row histograms   M[row][bin]      (reads rating+mode, 5 B/player)
//   k_colscan  column prefix + bin bases + per-segment lobby arithmetic (tiny)
//   k_place    stable rank inside the row -> final lobby-major slot; scatters
//              player_id straight to member_ids (reads 13 B/player, writes 8 B)
//   k_finish   residual players -> compacted pool (enqueue order kept)
//   k_headers  lobby headers from the segment table
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
constexpr uint32_t kResCap = 2048;    // residual players one row may hold
constexpr uint32_t kMaxRows = 2048;   // rows (CTAs) of the histogram matrix
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
  uint32_t ticket;
  uint32_t n_lobbies, n_matched, n_alive, n_dead, n_resid;
  uint32_t overflow;
  uint32_t pad;
};

struct ActiveView {
  unsigned long long* keys;
  unsigned long long* vals;
  uint64_t mask;  // capacity - 1, 0 = no active set
};

__device__ __forceinline__ uint64_t hash64(uint64_t x) {
  x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
  return x;
}


// In-place exclusive scan of a shared-memory array a[0..n) by the whole CTA; returns the
// total.  s_tmp must hold BLOCK words.  (n is a few hundred to a few thousand.)
template <int BLOCK>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t* a, uint32_t n, uint32_t* s_tmp) {
  const uint32_t tid = threadIdx.x;
  const uint32_t per = (n + BLOCK - 1) / BLOCK;
  const uint32_t lo = tid * per < n ? tid * per : n, hi = (lo + per < n) ? lo + per : n;
  uint32_t local = 0;
  for (uint32_t i = lo; i < hi; ++i) local += a[i];
  s_tmp[tid] = local;
  __syncthreads();
  for (int off = 1; off < BLOCK; off <<= 1) {
    const uint32_t v = (tid >= (uint32_t)off) ? s_tmp[tid - off] : 0;
    __syncthreads();
    s_tmp[tid] += v;
    __syncthreads();
  }
  uint32_t run = s_tmp[tid] - local;
  for (uint32_t i = lo; i < hi; ++i) { const uint32_t v = a[i]; a[i] = run; run += v; }
  const uint32_t total = s_tmp[BLOCK - 1];
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
// k_hist: M[row][bin] = number of the row's players in that bin.
// Coalesced 128-bit rating loads (4 players per thread), 32-bit mode loads.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_hist(PoolView p, uint32_t n, uint32_t chunk, BinMap bm, uint32_t Kp,
                                                 uint32_t* __restrict__ M) {
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
  for (uint32_t i = beg + tid * 4; i < end; i += kBlock * 4) {
    if (i + 4 <= end) {
      const int4 r = __ldcs(reinterpret_cast<const int4*>(p.rating + i));
      const uint32_t m = __ldcs(reinterpret_cast<const uint32_t*>(p.mode + i));
      atomicAdd(&hist[bin_of(bm, s_lut, r.x, m & 0xFF)], 1u);
      atomicAdd(&hist[bin_of(bm, s_lut, r.y, (m >> 8) & 0xFF)], 1u);
      atomicAdd(&hist[bin_of(bm, s_lut, r.z, (m >> 16) & 0xFF)], 1u);
      atomicAdd(&hist[bin_of(bm, s_lut, r.w, m >> 24)], 1u);
    } else {
      for (uint32_t e = i; e < end; ++e) atomicAdd(&hist[bin_of(bm, s_lut, p.rating[e], p.mode[e])], 1u);
    }
  }
  __syncthreads();
  uint32_t* row = M + (size_t)blockIdx.x * Kp;
  for (uint32_t i = tid; i < Kp; i += kBlock) row[i] = hist[i];
}

// ---------------------------------------------------------------------------------------
// k_colscan: exclusive prefix down every column of M; the last block to finish then
// scans the bin totals and does the per-segment lobby arithmetic:
//   lobbies_s = n_s / L,  matched_s = lobbies_s * L,  residual_s = n_s - matched_s
// outbase[bin] = sorted position of the bin's first player minus the residual players
// of earlier segments (= its slot in member_ids); binlim[bin] = end of the segment's
// matched slots.  A player whose slot is >= binlim stays queued.
// ---------------------------------------------------------------------------------------
constexpr int kScanBlock = 256;
constexpr uint32_t kMaxSegs = MM_MAX_GROUPS * MM_MAX_MODES;

__global__ void __launch_bounds__(kScanBlock) k_colscan(uint32_t R, uint32_t Kp, uint32_t K, uint32_t* __restrict__ M,
                                                        uint32_t* __restrict__ tot, uint32_t* __restrict__ binbase,
                                                        uint32_t* __restrict__ outbase, uint32_t* __restrict__ binlim,
                                                        const uint32_t* __restrict__ seg_bin_lo,
                                                        const uint32_t* __restrict__ seg_L, uint32_t n_segs,
                                                        SegInfo* __restrict__ seg, uint32_t* __restrict__ seg_shift,
                                                        TickCtr* ctr) {
  const uint32_t tid = threadIdx.x;
  const uint32_t b = blockIdx.x * kScanBlock + tid;
  if (b < Kp) {
    uint32_t run = 0;
    uint32_t* col = M + b;
    constexpr int kT = 16;  // rows loaded per step: keeps 16 independent loads in flight
    for (uint32_t r0 = 0; r0 < R; r0 += kT) {
      uint32_t v[kT];
#pragma unroll
      for (int t = 0; t < kT; ++t) v[t] = (r0 + t < R) ? __ldcg(col + (size_t)(r0 + t) * Kp) : 0u;
#pragma unroll
      for (int t = 0; t < kT; ++t) {
        if (r0 + t < R) col[(size_t)(r0 + t) * Kp] = run;
        run += v[t];
      }
    }
    tot[b] = run;
  }
  __shared__ uint32_t s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&ctr->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // exclusive scan of tot[0..Kp) -> binbase[0..Kp]
  __shared__ uint32_t s_sum[kScanBlock];
  const uint32_t per = (Kp + kScanBlock - 1) / kScanBlock;
  const uint32_t lo = tid * per, hi = (lo + per < Kp) ? lo + per : Kp;
  uint32_t local = 0;
  for (uint32_t i = lo; i < hi; ++i) local += __ldcg(&tot[i]);
  s_sum[tid] = local;
  __syncthreads();
  for (int off = 1; off < kScanBlock; off <<= 1) {  // Hillis-Steele inclusive scan
    uint32_t v = (tid >= (uint32_t)off) ? s_sum[tid - off] : 0;
    __syncthreads();
    s_sum[tid] += v;
    __syncthreads();
  }
  uint32_t run = s_sum[tid] - local;
  for (uint32_t i = lo; i < hi; ++i) { binbase[i] = run; run += __ldcg(&tot[i]); }
  if (tid == kScanBlock - 1) binbase[Kp] = s_sum[kScanBlock - 1];
  __threadfence_block();
  __syncthreads();

  // per-segment arithmetic: three small scans over the <= modes*groups segments
  __shared__ uint32_t s_res[kMaxSegs], s_lob[kMaxSegs], s_n[kMaxSegs];
  for (uint32_t s = tid; s < n_segs; s += kScanBlock) {
    const uint32_t ns = binbase[seg_bin_lo[s + 1]] - binbase[seg_bin_lo[s]];
    const uint32_t nl = ns / seg_L[s];
    s_n[s] = ns; s_lob[s] = nl; s_res[s] = ns - nl * seg_L[s];
  }
  __syncthreads();
  for (uint32_t s = tid; s < n_segs; s += kScanBlock) { seg[s].n = s_n[s]; seg[s].n_lobbies = s_lob[s]; }
  const uint32_t tot_res = block_excl_scan<kScanBlock>(s_res, n_segs, s_sum);
  const uint32_t tot_lob = block_excl_scan<kScanBlock>(s_lob, n_segs, s_sum);
  const uint32_t tot_alive = block_excl_scan<kScanBlock>(s_n, n_segs, s_sum);
  for (uint32_t s = tid; s < n_segs; s += kScanBlock) {
    seg[s].member_base = binbase[seg_bin_lo[s]] - s_res[s];
    seg[s].lobby_base = s_lob[s];
    seg_shift[s] = s_res[s];
  }
  if (tid == 0) {
    ctr->n_lobbies = tot_lob; ctr->n_matched = tot_alive - tot_res; ctr->n_alive = tot_alive;
    ctr->n_dead = __ldcg(&tot[K]);
  }
  __threadfence_block();
  __syncthreads();
  for (uint32_t i = tid; i < Kp; i += kScanBlock) {
    if (i >= K) { outbase[i] = 0; binlim[i] = 0; continue; }
    uint32_t a = 0, c = n_segs;  // last s with seg_bin_lo[s] <= i
    while (c - a > 1) { const uint32_t mid = (a + c) >> 1; if (seg_bin_lo[mid] <= i) a = mid; else c = mid; }
    // skip empty segments that share the same lower bound
    while (a + 1 < n_segs && seg_bin_lo[a + 1] <= i) ++a;
    outbase[i] = binbase[i] - seg_shift[a];
    binlim[i] = seg[a].member_base + seg[a].n_lobbies * seg_L[a];
  }
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
// Bit 31 of cnt marks a (row, bin) cell that reaches past the segment's matched range:
// only those players consult binlim (the < L leftovers of a partition stay queued).
// IMPL 0 is a slow warp-serial ranking kept as an on-device cross-check.
// ---------------------------------------------------------------------------------------
template <int IMPL>
__global__ void __launch_bounds__(kBlock, 1)
    k_place(PoolView p, uint32_t n, uint32_t chunk, BinMap bm, uint32_t Kp, uint32_t R, const uint32_t* __restrict__ M,
            const uint32_t* __restrict__ tot, const uint32_t* __restrict__ outbase, const uint32_t* __restrict__ binlim,
            uint64_t* __restrict__ members, uint32_t* __restrict__ src_idx, uint32_t* __restrict__ resid_stage,
            uint32_t* __restrict__ rescnt, TickCtr* ctr) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* cnt = smem;
  uint32_t* head = cnt + Kp;                                 // IMPL 1 only
  uint32_t* node = head + (IMPL == 1 ? Kp : 0);              // [kRound]
  uint32_t* res_list = node + (IMPL == 1 ? kRound : 0);      // [kResCap]
  uint16_t* s_lut = reinterpret_cast<uint16_t*>(res_list + kResCap);
  __shared__ uint32_t s_nres;

  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t row = blockIdx.x;
  {
    const uint32_t* mrow = M + (size_t)row * Kp;
    const uint32_t* mnext = (row + 1 < R) ? mrow + Kp : tot;  // prefix of the next row, or column total
    for (uint32_t i = tid; i < Kp; i += kBlock) {
      const uint32_t pre = mrow[i], c = mnext[i] - pre;
      const uint32_t start = outbase[i] + pre;
      const uint32_t flag = (start + c > binlim[i]) ? 0x80000000u : 0u;
      cnt[i] = start | flag;
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

  for (uint32_t round = 0; round < n_rounds; ++round) {
    const uint32_t base = beg + round * kRound;
    uint32_t bin[kJ];
    uint64_t idv[kJ];
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      const uint32_t e = base + j * kBlock + tid;
      if (e < end) {
        const int32_t r = __ldcs(p.rating + e);
        const uint32_t m = __ldcs(p.mode + e);
        idv[j] = __ldcs(reinterpret_cast<const unsigned long long*>(p.id + e));
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
      if (bin[j] < bm.K) {
        const uint32_t e = base + j * kBlock + tid;
        const uint32_t slot = (base_g[j] & 0x7FFFFFFFu) + rankw[j];
        bool matched = true;
        if (base_g[j] >> 31) matched = slot < __ldg(&binlim[bin[j]]);
        if (matched) {
          members[slot] = idv[j];
          if (src_idx) src_idx[slot] = e;
        } else {
          const uint32_t k = atomicAdd(&s_nres, 1u);
          if (k < kResCap) res_list[k] = e;
        }
      }
    }
    if (IMPL == 1) __syncthreads();
  }

  // the row's residual players, in enqueue order
  __syncthreads();
  const uint32_t nres_all = s_nres;
  const uint32_t nres = nres_all < kResCap ? nres_all : kResCap;
  if (tid == 0) {
    rescnt[row] = nres;
    if (nres_all > kResCap) atomicExch(&ctr->overflow, 1u);
  }
  for (uint32_t t = tid; t < nres; t += kBlock) {
    const uint32_t v = res_list[t];
    uint32_t rank = 0;
    for (uint32_t u = 0; u < nres; ++u) rank += (res_list[u] < v) ? 1u : 0u;
    resid_stage[(size_t)row * kResCap + rank] = v;
  }
}

// ---------------------------------------------------------------------------------------
// k_finish (one CTA): concatenate the rows' residual lists (rows are in enqueue order),
// gather the five pool columns into the alternate pool buffer and re-stamp the residual
// players' active-set entries with their new slot.  Replaces save_new_state/3
// (search/worker.ex:282-289): the "partial lobby" is simply the players left resident.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_finish(PoolView src, PoolView dst, uint32_t R, const uint32_t* __restrict__ rescnt,
                                                 const uint32_t* __restrict__ resid_stage,
                                                 ActiveView act, uint32_t new_gen, TickCtr* ctr) {
  __shared__ uint32_t s_off[kMaxRows + 1];
  __shared__ uint32_t s_tmp[1024];
  const uint32_t tid = threadIdx.x;
  for (uint32_t r = tid; r < R; r += 1024) s_off[r] = rescnt[r];
  __syncthreads();
  const uint32_t s_total = block_excl_scan<1024>(s_off, R, s_tmp);
  if (tid == 0) { s_off[R] = s_total; ctr->n_resid = s_total; }
  __syncthreads();
  const uint32_t total = s_total;
  for (uint32_t t = tid; t < total; t += 1024) {
    uint32_t a = 0, c = R;  // last row with s_off[row] <= t
    while (c - a > 1) { const uint32_t mid = (a + c) >> 1; if (s_off[mid] <= t) a = mid; else c = mid; }
    const uint32_t idx = resid_stage[(size_t)a * kResCap + (t - s_off[a])];
    const uint64_t pid = src.id[idx];
    dst.id[t] = pid; dst.rating[t] = src.rating[idx]; dst.mode[t] = src.mode[idx];
    dst.tsize[t] = src.tsize[idx]; dst.ts[t] = src.ts[idx];
    if (act.mask) {
      uint64_t h = hash64(pid) & act.mask;
      for (uint64_t probe = 0; probe <= act.mask; ++probe) {
        const unsigned long long k = act.keys[h];
        if (k == pid) { act.vals[h] = ((unsigned long long)new_gen << 32) | t; break; }
        if (k == kEmptyKey) break;
        h = (h + 1) & act.mask;
      }
    }
  }
}

// k_headers: lobby c of segment s = members [member_base + k*L, +L).  Replaces the
// payload assembly at search/worker.ex:315-319 ({"teams": ..., "game-mode": ...}).
__global__ void k_headers(const SegInfo* __restrict__ seg, const uint32_t* __restrict__ seg_L, uint32_t n_segs,
                          uint32_t n_groups, const TickCtr* __restrict__ ctr, mm_lobby_hdr* __restrict__ hdr,
                          const uint32_t* __restrict__ src_idx, uint32_t* __restrict__ emit_seq) {
  const uint32_t total = ctr->n_lobbies;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < total; c += gridDim.x * blockDim.x) {
    uint32_t a = 0, e = n_segs;
    while (e - a > 1) { const uint32_t mid = (a + e) >> 1; if (seg[mid].lobby_base <= c) a = mid; else e = mid; }
    while (a + 1 < n_segs && seg[a + 1].lobby_base <= c) ++a;
    const uint32_t L = seg_L[a];
    mm_lobby_hdr h;
    h.first_member = seg[a].member_base + (c - seg[a].lobby_base) * L;
    h.n_members = (uint16_t)L;
    h.mode = (uint8_t)(a / n_groups);
    h.group = (uint8_t)(a % n_groups);
    hdr[c] = h;
    if (emit_seq) emit_seq[c] = src_idx[h.first_member + L - 1];
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
__global__ void k_enq_claim(uint32_t n, const uint64_t* __restrict__ id, const int32_t* __restrict__ rating,
                            const uint8_t* __restrict__ mode, const uint8_t* __restrict__ grp_lut, int32_t key_lo,
                            uint32_t KR, uint32_t n_modes, ActiveView act, uint64_t* __restrict__ hslot,
                            uint8_t* __restrict__ code) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
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
__global__ void k_enq_count(uint32_t n, ActiveView act, const uint64_t* __restrict__ hslot, uint8_t* __restrict__ code,
                            uint32_t* __restrict__ blocksum) {
  __shared__ uint32_t s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool win = false;
  if (i < n && code[i] == 1) {
    win = !act.mask || act.vals[hslot[i]] == (kPending | i);
    if (!win) code[i] = 0;  // a lower batch index holds the id
  }
  const uint32_t b = __ballot_sync(0xFFFFFFFFu, win);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(&s_cnt, (uint32_t)__popc(b));
  __syncthreads();
  if (threadIdx.x == 0) blocksum[blockIdx.x] = s_cnt;
}

// exclusive scan of blocksum (single CTA; nblocks is at most a few 10k)
__global__ void __launch_bounds__(1024) k_scan_small(uint32_t nb, uint32_t* __restrict__ v, uint32_t* __restrict__ total) {
  __shared__ uint32_t s_sum[1024];
  const uint32_t tid = threadIdx.x;
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
  uint32_t run = s_sum[tid] - local;
  for (uint32_t i = lo; i < hi && i < nb; ++i) { const uint32_t x = v[i]; v[i] = run; run += x; }
  if (tid == 1023) *total = s_sum[1023];
}

// E3: append winners to the pool in batch order (= enqueue order) and commit their
// active-set entries.  Players past the pool capacity are rolled back with code 3.
__global__ void k_enq_append(uint32_t n, const uint64_t* __restrict__ id, const int32_t* __restrict__ rating,
                             const uint8_t* __restrict__ mode, const uint32_t* __restrict__ ts,
                             const uint8_t* __restrict__ mode_tsize, ActiveView act, const uint64_t* __restrict__ hslot,
                             uint8_t* __restrict__ code, const uint32_t* __restrict__ blockoff, PoolView pool,
                             uint32_t n_pool, uint32_t capacity, uint32_t gen, uint32_t* __restrict__ n_rejected_cap) {
  __shared__ uint32_t s_warp[32];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool win = i < n && code[i] == 1;
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
  if (act.mask) act.vals[hslot[i]] = ((unsigned long long)gen << 32) | slot;
}

// ActiveUser.remove_user/1 (models/active_user.ex:57-66), batched.  A player still
// queued is tombstoned in the pool (mode byte = DEAD) so the next tick drops it the way
// remove_inactive_players/1 (search/worker.ex:267-280) filters it.
__global__ void k_remove(uint32_t n, const uint64_t* __restrict__ id, ActiveView act, PoolView pool, uint32_t n_pool,
                         uint32_t gen, uint32_t* __restrict__ n_removed) {
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
      if (v < kPending && g == gen && slot < n_pool && pool.id[slot] == pid) pool.mode[slot] = MM_MODE_DEAD;
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

__global__ void k_fill64(unsigned long long* p, uint64_t n, unsigned long long v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace mm
