// mm_epilogue.cuh — phase 4 of the tick: pool compaction of the leftovers + lobby headers
#pragma once
#include "mm_common.cuh"
#include "mm_scan.cuh"

namespace mm {

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

}  // namespace mm
