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
#include "mm_common.cuh"
#include "mm_hist.cuh"
#include "mm_scan.cuh"
#include "mm_place.cuh"
#include "mm_epilogue.cuh"
#include "mm_active.cuh"

namespace mm {

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

}  // namespace mm
