// mm_active.cuh — active set (GPU hash table) and pool ingest kernels
#pragma once
#include "mm_common.cuh"

namespace mm {

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
