// mm_engine.cu — C ABI of the B200 opponent-search engine (include/mm_engine.h).
//
// Host side of the drop-in for the reference search stage
// (matchmaking/lib/search/worker.ex + models/{active_user,lobby_state}.ex).  The pool
// is a GPU-resident SoA (player_id u64 / rating i32 / game-mode u8 / team-size u8 /
// enqueue-time u32) kept in enqueue order; all matching work runs in the kernels of
// mm_kernels.cuh.  There is no CPU path: every entry point either launches CUDA work
// or fails with MM_E_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "mm_kernels.cuh"

using namespace mm;

namespace {

struct Pool {
  PoolView v{};
  uint32_t n = 0;
};

struct Table {
  unsigned long long* kv = nullptr;  // hcap x {key, value}
};

}  // namespace

struct mm_engine {
  mm_config cfg{};
  std::mutex mu;
  int device = 0;
  int n_sms = 0;
  size_t smem_optin = 0, smem_sm = 0;
  int block = 1024;  // threads per row CTA (512 when two rows share an SM)
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cudaStream_t copy_stream = nullptr;  // H2D of ingest chunks, overlapped with the claim kernels
  cudaEvent_t ev_copy = nullptr;
  cudaStream_t d2h_stream = nullptr;   // async_results: a tick's host copies, overlapped with the next ingest
  bool async_results = false, results_pending = false;
  cudaEvent_t ev[5]{};  // tick start | after hist | after colscan | after place | after epilogue
  char last_err[512] = {0};

  // key domain
  int32_t key_lo = 0;
  uint32_t KR = 0, stride = 0, K = 0, Kp = 0, n_segs = 0;
  uint16_t* d_lut = nullptr;
  uint8_t* d_grp_lut = nullptr;
  uint8_t* d_mode_tsize = nullptr;
  uint32_t* d_seg_bin_lo = nullptr;
  uint32_t* d_seg_L = nullptr;
  uint16_t* d_bin_seg = nullptr;  // [Kp] bin -> (mode, group) segment
  uint32_t min_L = 1;

  // pool (double buffered) + snapshot
  uint32_t capacity = 0;
  Pool pool[2];
  int cur = 0;
  uint32_t gen = 1;
  Pool snap;
  uint32_t snap_gen = 0;
  bool has_snap = false;

  // active set
  bool use_active = true;
  uint64_t hcap = 0;
  Table tab[2];
  int tcur = 0;
  uint64_t n_active = 0, n_tomb = 0;

  // tick scratch
  uint32_t R = 0;
  int rows_per_sm = 1;
  int rank_impl = 1;
  int place_debug = 0;
  size_t persist_bytes = 0;
  uint32_t place2_stages = 0;  // 0 = k_place2 does not fit in shared memory
  uint32_t hist3_stages = 4;   // ring depth of the bin-column histogram
  int fused_ok = 0;  // k_tick<512> can be launched cooperatively with 2 CTAs per SM
  int tick_impl = 1; // 1 = one fused cooperative launch when possible, 0 = four launches
  size_t tick_smem = 0;
  int dense_ok = 2;            // small-K ranking: 0 = off, 1 = MATCH-based matrix, 2 = private byte counters when possible
  uint32_t* d_M = nullptr;
  uint32_t *d_tot = nullptr, *d_outbase = nullptr, *d_binlim = nullptr;
  uint16_t* d_bin_key = nullptr;
  int32_t max_spread = -1;  // < 0: policy S0 (reference behaviour); >= 0: policy S1 (extension)
  SegInfo* d_seg = nullptr;
  uint32_t* d_left_bits = nullptr;  // one bit per pool slot: stays queued after the tick
  uint64_t* d_members = nullptr;
  uint32_t* d_src_idx = nullptr;
  mm_lobby_hdr* d_hdr = nullptr;
  uint32_t* d_emit_seq = nullptr;
  uint32_t max_lobbies = 0;
  uint32_t* d_rescnt = nullptr;
  TickCtr* d_ctr = nullptr;
  TickCtr* h_ctr = nullptr;  // pinned

  // enqueue scratch (grown on demand)
  uint32_t enq_cap = 0;
  uint64_t *d_in_id = nullptr, *d_hslot = nullptr;
  int32_t* d_in_rating = nullptr;
  uint8_t *d_in_mode = nullptr, *d_code = nullptr;
  uint32_t *d_in_ts = nullptr, *d_blocksum = nullptr, *d_small = nullptr;  // d_small: [0]=total [1]=rejected [2]=removed
  uint32_t* h_small = nullptr;                                             // pinned

  // last tick
  mm_tick_stats last{};
  bool last_fused = false;
};

namespace {

int fail(mm_engine* e, cudaError_t err, const char* what) {
  if (e) std::snprintf(e->last_err, sizeof(e->last_err), "%s: %s", what, cudaGetErrorString(err));
  cudaGetLastError();  // clear sticky-free errors
  return MM_E_CUDA;
}
#define CK(call)                                          \
  do {                                                    \
    cudaError_t _err = (call);                            \
    if (_err != cudaSuccess) return fail(e, _err, #call); \
  } while (0)

// Allow a kernel the device's whole opt-in shared memory (minus its static part).  Function attributes are
// process-global: an engine with a small key domain must never lower the limit another engine relies on.
template <class F>
cudaError_t allow_max_smem(const mm_engine* e, F* func) {
  cudaFuncAttributes fa{};
  cudaError_t err = cudaFuncGetAttributes(&fa, func);
  if (err != cudaSuccess) return err;
  return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(e->smem_optin - fa.sharedSizeBytes));
}

int alloc_pool(mm_engine* e, Pool& p, uint32_t cap) {
  const size_t c = (size_t)cap + 3 * kRound;  // TMA tiles are read whole: pad past the last row
  CK(cudaMalloc(&p.v.id, c * 8));
  CK(cudaMalloc(&p.v.rating, c * 4));
  CK(cudaMalloc(&p.v.mode, c));
  CK(cudaMalloc(&p.v.tsize, c));
  CK(cudaMalloc(&p.v.ts, c * 4));
  CK(cudaMalloc(&p.v.bin, c * 2));
  p.n = 0;
  return MM_OK;
}
void free_pool(Pool& p) {
  cudaFree(p.v.id); cudaFree(p.v.rating); cudaFree(p.v.mode); cudaFree(p.v.tsize); cudaFree(p.v.ts); cudaFree(p.v.bin);
  p = Pool{};
}

ActiveView act_view(mm_engine* e) {
  ActiveView a{};
  if (e->use_active) { a.keys.p = e->tab[e->tcur].kv; a.vals.p = e->tab[e->tcur].kv + 1; a.mask = e->hcap - 1; }
  return a;
}

int clear_table(mm_engine* e, Table& t) {
  k_fill_kv<<<1024, 256, 0, e->stream>>>(reinterpret_cast<ulonglong2*>(t.kv), e->hcap, kEmptyKey, kFreeVal);
  CK(cudaGetLastError());
  return MM_OK;
}

// generic/worker.ex:46-53 on the host (also exported as mm_group_of)
int group_of(const mm_config* cfg, int64_t rating) {
  for (uint32_t g = 0; g < cfg->n_groups; ++g)
    if (rating >= cfg->group_lo[g] && rating <= cfg->group_hi[g]) return (int)g;
  return cfg->default_group;
}

int check_config(const mm_config* c) {
  if (!c || c->abi_version != MM_ABI_VERSION) return MM_E_ARG;
  if (c->n_groups == 0 || c->n_groups > MM_MAX_GROUPS || c->n_modes == 0 || c->n_modes > MM_MAX_MODES) return MM_E_ARG;
  if (c->default_group >= (int32_t)c->n_groups || c->default_group < -1) return MM_E_ARG;
  if (c->order_mode > MM_ORDER_RATING) return MM_E_ARG;
  if (c->capacity == 0 || c->capacity > 0x7FFF0000u) return MM_E_ARG;
  for (uint32_t g = 0; g < c->n_groups; ++g) {
    if (c->group_lo[g] > c->group_hi[g]) return MM_E_ARG;
    if (c->group_lo[g] < -(1 << 30) || c->group_hi[g] > (1 << 30)) return MM_E_ARG;
  }
  for (uint32_t m = 0; m < c->n_modes; ++m) {
    const uint32_t L = (uint32_t)c->modes[m].teams * c->modes[m].team_size;
    if (L == 0 || L > 65535u || c->modes[m].team_size > 255) return MM_E_ARG;
  }
  return MM_OK;
}

size_t place_smem(const mm_engine* e, int impl) {
  size_t words = e->Kp + (impl == 1 ? (size_t)e->Kp + kRound : 0);
  return words * 4 + (size_t)e->KR * 2 + 16;
}
bool place2_dense(const mm_engine* e) { return e->Kp <= kDenseMaxBins; }
// private byte counters c8[Kp][512] + lane bases + bases: only for very few bins and 512-thread CTAs
bool place2_dense2(const mm_engine* e) { return e->Kp <= 96; }
size_t place2_dense_bytes(const mm_engine* e) {
  size_t a = place2_dense(e) ? (size_t)e->Kp * kDenseStride * 2 * 2 + (size_t)e->Kp * 4 : 0;
  size_t b = place2_dense2(e) ? (size_t)e->Kp * (512 + 16) + (size_t)e->Kp * 68 + (size_t)e->Kp * 4 + 32 : 0;
  return std::max(a, b);
}
size_t place2_smem(const mm_engine* e, uint32_t stages) {
  return (size_t)stages * kTileBytes + 64 + ((size_t)e->Kp + kHeadSlots + kTile) * 4 + (size_t)kTile * 2 +
         place2_dense_bytes(e) + 16;
}
// Shared-memory layout of the scan tail: everything on chip up to ~100 KB (so that it never exceeds the placement
// phase's footprint in the fused kernel), else keys from global memory, else matched counts parked in global too.
uint32_t tail_layout(const mm_engine* e) {
  if ((size_t)tail_words(e->Kp, 3) * 4 <= 100 * 1024) return 3;
  if ((size_t)tail_words(e->Kp, 1) * 4 + 1024 <= e->smem_optin) return 1;
  return 0;
}
size_t colscan_smem(const mm_engine* e) {
  return (size_t)std::max<uint32_t>(kColScratchWords, tail_words(e->Kp, tail_layout(e))) * 4;
}
size_t hist3_smem(const mm_engine* e, uint32_t stages) { return (size_t)stages * kBTileBytes + 64 + (size_t)e->Kp * 4 + 16; }
size_t hist_smem(const mm_engine* e) { return (size_t)e->Kp * 4 + (size_t)e->KR * 2 + 16; }

// Build the key -> bin LUT and the (mode, group) segment table (see mm_kernels.cuh).
int build_tables(mm_engine* e) {
  const mm_config& c = e->cfg;
  int32_t rmin = c.group_lo[0], rmax = c.group_hi[0];
  for (uint32_t g = 1; g < c.n_groups; ++g) { rmin = std::min(rmin, c.group_lo[g]); rmax = std::max(rmax, c.group_hi[g]); }
  e->key_lo = rmin - 1;
  const uint64_t KR64 = (uint64_t)((int64_t)rmax - rmin + 3);
  if (KR64 > 65535u) return MM_E_ARG;  // rating span must fit the 16-bit LUT
  e->KR = (uint32_t)KR64;
  const uint32_t G = c.n_groups;
  std::vector<uint8_t> grp(e->KR);
  for (uint32_t k = 0; k < e->KR; ++k) {
    const int g = group_of(&c, (int64_t)e->key_lo + k);
    grp[k] = g < 0 ? 0xFF : (uint8_t)g;
  }
  std::vector<uint16_t> lut(e->KR, 0);
  std::vector<uint32_t> first(G + 1, 0);
  std::vector<uint16_t> key_of;  // rating order: bin (inside a mode) -> clamp key
  if (c.order_mode == MM_ORDER_RATING) {
    // bins ordered by (group, clamp key): the partition of a group is its keys ascending
    uint32_t next = 0;
    for (uint32_t g = 0; g < G; ++g) {
      first[g] = next;
      for (uint32_t k = 0; k < e->KR; ++k)
        if (grp[k] == g) { lut[k] = (uint16_t)next++; key_of.push_back((uint16_t)k); }
    }
    first[G] = next;
    e->stride = std::max(next, 1u);
  } else {
    for (uint32_t k = 0; k < e->KR; ++k) lut[k] = grp[k] == 0xFF ? 0 : grp[k];
    for (uint32_t g = 0; g <= G; ++g) first[g] = g;
    e->stride = G;
  }
  e->K = c.n_modes * e->stride;
  e->Kp = e->K + 1;
  e->n_segs = c.n_modes * G;
  std::vector<uint32_t> seg_lo(e->n_segs + 1), seg_L(e->n_segs);
  e->min_L = 0xFFFFFFFFu;
  std::vector<uint8_t> tsz(MM_MAX_MODES, 0);
  for (uint32_t m = 0; m < c.n_modes; ++m) {
    const uint32_t L = (uint32_t)c.modes[m].teams * c.modes[m].team_size;
    e->min_L = std::min(e->min_L, L);
    tsz[m] = (uint8_t)c.modes[m].team_size;
    for (uint32_t g = 0; g < G; ++g) { seg_lo[m * G + g] = m * e->stride + first[g]; seg_L[m * G + g] = L; }
  }
  seg_lo[e->n_segs] = e->K;
  std::vector<uint16_t> bin_seg(e->Kp, 0);
  for (uint32_t sgi = 0; sgi < e->n_segs; ++sgi)
    for (uint32_t b = seg_lo[sgi]; b < seg_lo[sgi + 1]; ++b) bin_seg[b] = (uint16_t)sgi;
  std::vector<uint16_t> bin_key(e->Kp, 0);
  if (!key_of.empty())
    for (uint32_t b = 0; b < e->K; ++b) bin_key[b] = key_of[b % e->stride];
  CK(cudaMalloc(&e->d_bin_seg, e->Kp * 2));
  CK(cudaMemcpy(e->d_bin_seg, bin_seg.data(), e->Kp * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&e->d_bin_key, e->Kp * 2));
  CK(cudaMemcpy(e->d_bin_key, bin_key.data(), e->Kp * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&e->d_lut, e->KR * 2));
  CK(cudaMalloc(&e->d_grp_lut, e->KR));
  CK(cudaMalloc(&e->d_mode_tsize, MM_MAX_MODES));
  CK(cudaMalloc(&e->d_seg_bin_lo, (e->n_segs + 1) * 4));
  CK(cudaMalloc(&e->d_seg_L, e->n_segs * 4));
  CK(cudaMemcpy(e->d_lut, lut.data(), e->KR * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->d_grp_lut, grp.data(), e->KR, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->d_mode_tsize, tsz.data(), MM_MAX_MODES, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->d_seg_bin_lo, seg_lo.data(), (e->n_segs + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(e->d_seg_L, seg_L.data(), e->n_segs * 4, cudaMemcpyHostToDevice));
  return MM_OK;
}

BinMap bin_map(const mm_engine* e) {
  BinMap b{};
  b.lut = e->d_lut; b.key_lo = e->key_lo; b.KR = e->KR; b.stride = e->stride; b.K = e->K;
  return b;
}

int alloc_tick_scratch(mm_engine* e) {
  e->R = (uint32_t)e->n_sms * (uint32_t)e->rows_per_sm;
  if (e->R > kMaxRows) e->R = kMaxRows;
  if (e->d_M) { cudaFree(e->d_M); cudaFree(e->d_rescnt); }
  CK(cudaMalloc(&e->d_M, (size_t)e->R * e->Kp * 4));
  CK(cudaMalloc(&e->d_rescnt, (size_t)(e->R + 1) * 4));
  return MM_OK;
}

int ensure_enq_scratch(mm_engine* e, uint32_t n) {
  if (n <= e->enq_cap) return MM_OK;
  cudaFree(e->d_in_id); cudaFree(e->d_hslot); cudaFree(e->d_in_rating); cudaFree(e->d_in_mode);
  cudaFree(e->d_code); cudaFree(e->d_in_ts); cudaFree(e->d_blocksum);
  e->enq_cap = 0;
  const size_t c = (size_t)n + 64;
  CK(cudaMalloc(&e->d_in_id, c * 8));
  CK(cudaMalloc(&e->d_hslot, c * 8));
  CK(cudaMalloc(&e->d_in_rating, c * 4));
  CK(cudaMalloc(&e->d_in_mode, c));
  CK(cudaMalloc(&e->d_code, c));
  CK(cudaMalloc(&e->d_in_ts, c * 4));
  CK(cudaMalloc(&e->d_blocksum, (c / 256 + 2) * 4));
  e->enq_cap = n;
  return MM_OK;
}

// Drop tombstones: re-insert the committed entries into the spare table.
int rehash(mm_engine* e) {
  Table& nt = e->tab[e->tcur ^ 1];
  int rc = clear_table(e, nt);
  if (rc) return rc;
  ActiveView oldv = act_view(e), newv{{nt.kv}, {nt.kv + 1}, e->hcap - 1};
  k_rehash<<<2048, 256, 0, e->stream>>>(oldv, newv);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  e->tcur ^= 1;
  e->n_tomb = 0;
  return MM_OK;
}

// capacity checks + scratch for an ingest batch of n players
int enq_prepare(mm_engine* e, uint32_t n) {
  int rc = ensure_enq_scratch(e, n);
  if (rc) return rc;
  if (e->use_active && (e->n_active + e->n_tomb + n) * 4 > e->hcap * 3) {
    if ((e->n_active + n) * 4 > e->hcap * 3) return MM_E_CAP;  // active set full
    if ((rc = rehash(e))) return rc;
  }
  CK(cudaMemsetAsync(e->d_small, 0, 16, e->stream));
  return MM_OK;
}

// E1 on batch indices [base, base + cnt): validate + claim in the active set
int enq_claim(mm_engine* e, uint32_t base, uint32_t cnt, const uint64_t* id, const int32_t* rating, const uint8_t* mode) {
  k_enq_claim<<<(cnt + 255) / 256, 256, 0, e->stream>>>(base, cnt, id, rating, mode, e->d_grp_lut, e->key_lo, e->KR,
                                                         e->cfg.n_modes, act_view(e), e->d_hslot, e->d_code);
  CK(cudaGetLastError());
  return MM_OK;
}

// E2 + E3 on batch indices [base, base + cnt): winners, stable append to the pool (after the winners of the
// earlier chunks: the running total lives in d_small[0]), commit
int enq_append(mm_engine* e, uint32_t base, uint32_t cnt, const uint64_t* id, const int32_t* rating, const uint8_t* mode,
               const uint32_t* ts) {
  const uint32_t nb = (cnt + 255) / 256;
  Pool& p = e->pool[e->cur];
  ActiveView av = act_view(e);
  k_enq_count<<<nb, 256, 0, e->stream>>>(base, cnt, av, e->d_hslot, e->d_code, e->d_blocksum);
  k_scan_small<<<1, 1024, 0, e->stream>>>(nb, e->d_blocksum, e->d_small);
  k_enq_append<<<nb, 256, 0, e->stream>>>(base, cnt, id, rating, mode, ts, e->d_mode_tsize, av, e->d_hslot, e->d_code,
                                          e->d_blocksum, p.v, p.n, e->capacity, e->gen, e->d_small + 1, bin_map(e));
  CK(cudaGetLastError());
  return MM_OK;
}

// counters of the finished batch -> host state
int enq_finish(mm_engine* e, uint32_t n, uint8_t* accepted_dev, uint32_t* n_accepted) {
  Pool& p = e->pool[e->cur];
  CK(cudaMemcpyAsync(e->h_small, e->d_small, 16, cudaMemcpyDeviceToHost, e->stream));
  if (accepted_dev) CK(cudaMemcpyAsync(accepted_dev, e->d_code, n, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const uint32_t won = e->h_small[0], rej = e->h_small[1];
  const uint32_t acc = won - rej;
  p.n += acc;
  if (e->use_active) { e->n_active += acc; e->n_tomb += rej; }
  if (n_accepted) *n_accepted = acc;
  return MM_OK;
}

int enqueue_device_locked(mm_engine* e, uint32_t n, const uint64_t* id, const int32_t* rating, const uint8_t* mode,
                          const uint32_t* ts, uint8_t* accepted, uint32_t* n_accepted) {
  if (n_accepted) *n_accepted = 0;
  if (n == 0) return MM_OK;
  int rc = enq_prepare(e, n);
  if (rc) return rc;
  if ((rc = enq_claim(e, 0, n, id, rating, mode))) return rc;
  if ((rc = enq_append(e, 0, n, id, rating, mode, ts))) return rc;
  return enq_finish(e, n, accepted, n_accepted);
}

// Pin member_ids in a persisting L2 carve-out: the 8-byte scatter of k_place completes
// 32-byte sectors at unrelated times, so the lines must survive in L2 until the kernel ends
// (measured: -50 us and -79 MB of DRAM fill reads on the 10M-player tick).
int set_persist(mm_engine* e, int64_t mb) {
  CK(cudaStreamSynchronize(e->stream));
  cudaStreamAttrValue attr{};
  if (mb > 0) {
    int max_persist = 0, max_win = 0;
    CK(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, e->device));
    CK(cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, e->device));
    const size_t want = std::min<size_t>((size_t)mb << 20, (size_t)max_persist);
    if (want == 0) return MM_OK;
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    const size_t win = std::min<size_t>({(size_t)e->capacity * 8, (size_t)max_win});
    attr.accessPolicyWindow.base_ptr = e->d_members;
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = win ? std::min(1.0f, (float)want / (float)win) : 0.f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    e->persist_bytes = want;
  } else {
    attr.accessPolicyWindow.num_bytes = 0;
    e->persist_bytes = 0;
  }
  CK(cudaStreamSetAttribute(e->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
  return MM_OK;
}

uint32_t dense_mode(const mm_engine* e) {
  if (!e->dense_ok) return 0u;
  if (place2_dense2(e) && e->block == 512 && e->dense_ok != 1) return 2u;  // dense_ok: 1 = MATCH variant only
  return place2_dense(e) ? 1u : 0u;
}

TailArgs tail_args(const mm_engine* e) {
  TailArgs t{};
  t.Kp = e->Kp; t.K = e->K; t.n_segs = e->n_segs; t.max_spread = e->max_spread; t.layout = tail_layout(e);
  t.tot = e->d_tot; t.seg_bin_lo = e->d_seg_bin_lo; t.seg_L = e->d_seg_L; t.bin_seg = e->d_bin_seg;
  t.bin_key = e->d_bin_key; t.outbase = e->d_outbase; t.binlim = e->d_binlim; t.seg = e->d_seg; t.ctr = e->d_ctr;
  return t;
}

// launches k_hist + k_colscan and returns the counters (phase A of a tick)
int tick_phase_a(mm_engine* e, uint32_t n, uint32_t* chunk_out) {
  const Pool& p = e->pool[e->cur];
  uint32_t chunk = (n + e->R - 1) / e->R;
  chunk = std::max<uint32_t>(((chunk + kRound - 1) / kRound) * kRound, kRound);
  *chunk_out = chunk;
  CK(cudaMemsetAsync(e->d_ctr, 0, sizeof(TickCtr), e->stream));
  CK(cudaEventRecord(e->ev[0], e->stream));
  if (e->rank_impl == 3) {
    if (e->block == 512)
      k_hist3<512><<<e->R, 512, hist3_smem(e, e->hist3_stages), e->stream>>>(p.v.bin, n, chunk, e->Kp, e->hist3_stages, e->d_M,
                                                                              e->d_tot);
    else
      k_hist3<1024><<<e->R, 1024, hist3_smem(e, e->hist3_stages), e->stream>>>(p.v.bin, n, chunk, e->Kp, e->hist3_stages,
                                                                                e->d_M, e->d_tot);
  } else if (e->block == 512) {
    k_hist<512><<<e->R, 512, hist_smem(e), e->stream>>>(p.v, n, chunk, bin_map(e), e->Kp, e->d_M, e->d_tot, nullptr);
  } else {
    k_hist<1024><<<e->R, 1024, hist_smem(e), e->stream>>>(p.v, n, chunk, bin_map(e), e->Kp, e->d_M, e->d_tot, nullptr);
  }
  CK(cudaEventRecord(e->ev[1], e->stream));
  k_colscan<<<(e->Kp + 31) / 32 + 1, kScanBlock, colscan_smem(e), e->stream>>>(e->R, e->d_M, tail_args(e));
  CK(cudaGetLastError());
  return MM_OK;
}

int tick_phase_b(mm_engine* e, uint32_t n, uint32_t chunk, bool want_seq) {
  const Pool& p = e->pool[e->cur];
  Pool& q = e->pool[e->cur ^ 1];
  uint32_t* src_idx = want_seq ? e->d_src_idx : nullptr;
  CK(cudaEventRecord(e->ev[2], e->stream));
#define MM_PLACE(IMPL)                                                                                               \
  k_place<IMPL><<<e->R, kBlock, place_smem(e, IMPL), e->stream>>>(                                                    \
      p.v, n, chunk, bin_map(e), e->Kp, e->R, e->d_M, e->d_tot, e->d_outbase, e->d_binlim, e->d_members, src_idx,     \
      e->d_left_bits, e->d_rescnt, e->d_ctr)
#define MM_PLACE2(BLK)                                                                                               \
  k_place2<BLK><<<e->R, BLK, place2_smem(e, e->place2_stages), e->stream>>>(                                          \
      p.v.bin, p.v.id, n, chunk, e->K, e->Kp, e->R, e->place2_stages, dense_mode(e), e->d_M, e->d_tot,                \
      e->d_outbase, e->d_binlim, e->d_members, src_idx, e->d_left_bits, e->d_rescnt, e->d_ctr,                        \
      (uint32_t)e->place_debug)
  if (e->rank_impl == 3) {
    if (e->block == 512) MM_PLACE2(512);
    else MM_PLACE2(1024);
  } else if (e->rank_impl == 0) {
    MM_PLACE(0);
  } else {
    MM_PLACE(1);
  }
#undef MM_PLACE2
#undef MM_PLACE
  CK(cudaEventRecord(e->ev[3], e->stream));
  k_epilogue<<<std::max(1, e->n_sms), 1024, 0, e->stream>>>(p.v, q.v, n, chunk, e->R, e->d_rescnt, e->d_left_bits, act_view(e),
                                                           e->gen + 1, e->d_seg, e->d_seg_L, e->n_segs, e->cfg.n_groups,
                                                           e->d_hdr, src_idx, want_seq ? e->d_emit_seq : nullptr, e->d_tot,
                                                           e->Kp, e->d_ctr);
  CK(cudaGetLastError());
  CK(cudaEventRecord(e->ev[4], e->stream));
  CK(cudaMemcpyAsync(e->h_ctr, e->d_ctr, sizeof(TickCtr), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return MM_OK;
}

bool use_fused(const mm_engine* e) {
  return e->tick_impl == 1 && e->fused_ok && e->rank_impl == 3 && e->block == 512 && e->rows_per_sm == 2;
}

// the whole tick in one cooperative launch (k_tick)
int tick_fused(mm_engine* e, uint32_t n, bool want_seq) {
  const Pool& p = e->pool[e->cur];
  Pool& q = e->pool[e->cur ^ 1];
  uint32_t chunk = (n + e->R - 1) / e->R;
  chunk = std::max<uint32_t>(((chunk + kRound - 1) / kRound) * kRound, kRound);
  TickArgs a{};
  a.src = p.v; a.dst = q.v; a.left_bits = e->d_left_bits;
  a.n = n; a.chunk = chunk; a.R = e->R; a.n_groups = e->cfg.n_groups;
  a.hist_stages = e->hist3_stages; a.place_stages = e->place2_stages;
  a.dense = dense_mode(e);
  a.new_gen = e->gen + 1; a.dbg = (uint32_t)e->place_debug;
  a.M = e->d_M; a.tot = e->d_tot; a.tail = tail_args(e);
  a.members = e->d_members; a.src_idx = want_seq ? e->d_src_idx : nullptr; a.hdr = e->d_hdr;
  a.emit_seq = want_seq ? e->d_emit_seq : nullptr; a.rescnt = e->d_rescnt;
  a.act = act_view(e);
  CK(cudaMemsetAsync(e->d_ctr, 0, sizeof(TickCtr), e->stream));
  CK(cudaEventRecord(e->ev[0], e->stream));
  void* params[] = {&a};
  CK(cudaLaunchCooperativeKernel((const void*)k_tick<512>, dim3(e->R), dim3(512), params, e->tick_smem, e->stream));
  CK(cudaEventRecord(e->ev[4], e->stream));
  CK(cudaMemcpyAsync(e->h_ctr, e->d_ctr, sizeof(TickCtr), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return MM_OK;
}

int tick_commit(mm_engine* e, uint32_t n, mm_tick_stats* stats) {
  const TickCtr& c = *e->h_ctr;
  mm_tick_stats st{};
  st.pool_before = n; st.n_lobbies = c.n_lobbies; st.n_matched = c.n_matched; st.n_residual = c.n_resid;
  st.n_dead = c.n_dead; st.n_launches = 4;
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e->ev[0], e->ev[4]));
  st.device_us = ms * 1000.f;
  if (e->last_fused) {  // one launch: phase times from %globaltimer stamps of CTA 0
    st.n_launches = 1;
    st.hist_us = (float)(c.t[1] - c.t[0]) * 1e-3f;
    st.scan_us = (float)(c.t[2] - c.t[1]) * 1e-3f;
    st.place_us = (float)(c.t[3] - c.t[2]) * 1e-3f;
    st.epilogue_us = (float)(c.t[6] - c.t[3]) * 1e-3f;  // until the last CTA is done
    // debug: lobby headers done (max over CTAs) after barrier 3, in 10 ns units
    st.reserved = (uint32_t)((c.t[7] - c.t[3]) / 10);
    e->cur ^= 1;
    e->pool[e->cur].n = c.n_resid;
    e->gen += 1;
    e->last = st;
    if (stats) *stats = st;
    return MM_OK;
  }
  CK(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
  st.hist_us = ms * 1000.f;
  CK(cudaEventElapsedTime(&ms, e->ev[1], e->ev[2]));
  st.scan_us = ms * 1000.f;
  CK(cudaEventElapsedTime(&ms, e->ev[2], e->ev[3]));
  st.place_us = ms * 1000.f;
  CK(cudaEventElapsedTime(&ms, e->ev[3], e->ev[4]));
  st.epilogue_us = ms * 1000.f;
  e->cur ^= 1;
  e->pool[e->cur].n = c.n_resid;
  e->gen += 1;
  e->last = st;
  if (stats) *stats = st;
  return MM_OK;
}

// async_results: the previous tick's host copies must land before its device buffers are overwritten
int wait_results(mm_engine* e) {
  if (e->results_pending) {
    CK(cudaStreamSynchronize(e->d2h_stream));
    e->results_pending = false;
  }
  return MM_OK;
}

}  // namespace

// =======================================================================================
extern "C" {

uint32_t mm_abi_version(void) { return MM_ABI_VERSION; }

const char* mm_strerror(int s) {
  switch (s) {
    case MM_OK: return "ok";
    case MM_E_ARG: return "bad argument or config";
    case MM_E_CUDA: return "CUDA error or no CUDA device (no CPU fallback exists)";
    case MM_E_CAP: return "capacity exceeded";
    case MM_E_NCCL: return "NCCL error";
    case MM_E_STATE: return "invalid state for this call";
    default: return "unknown status";
  }
}

const char* mm_last_error(mm_engine* e) { return e ? e->last_err : ""; }

void mm_config_default(mm_config* c) {
  if (!c) return;
  std::memset(c, 0, sizeof(*c));
  c->abi_version = MM_ABI_VERSION;
  static const int32_t lo[7] = {0, 1500, 2000, 2500, 3000, 3500, 4000};  // config/config.exs:27-36
  static const int32_t hi[7] = {1499, 1999, 2499, 2999, 3499, 3999, 5000};
  c->n_groups = 7;
  for (int g = 0; g < 7; ++g) { c->group_lo[g] = lo[g]; c->group_hi[g] = hi[g]; }
  c->default_group = 7 / 2 + 1;  // generic/worker.ex:27 -> "diamond"
  c->n_modes = 2;
  c->modes[0].teams = 2; c->modes[0].team_size = 1;  // "1v1"
  c->modes[1].teams = 2; c->modes[1].team_size = 5;  // "5v5"
  c->order_mode = MM_ORDER_ARRIVAL;
  c->capacity = 1u << 20;
  c->active_capacity = 0;
  c->device = 0;
  c->flags = 0;
}

int mm_group_of(const mm_config* cfg, int32_t rating) {
  if (!cfg || cfg->n_groups == 0 || cfg->n_groups > MM_MAX_GROUPS) return MM_E_ARG;
  return group_of(cfg, rating);
}

int mm_create(const mm_config* cfg, mm_engine** out) {
  if (!out) return MM_E_ARG;
  *out = nullptr;
  int rc = check_config(cfg);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
    cudaGetLastError();
    return MM_E_CUDA;  // no CPU fallback
  }
  mm_engine* e = new (std::nothrow) mm_engine();
  if (!e) return MM_E_CAP;
  e->cfg = *cfg;
  e->device = cfg->device;
  e->capacity = cfg->capacity;
  e->use_active = !(cfg->flags & MM_F_NO_DEDUPE);
  auto bail = [&](int code) { mm_destroy(e); return code; };
  if (cudaSetDevice(e->device) != cudaSuccess) return bail(MM_E_CUDA);
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, e->device) != cudaSuccess) return bail(MM_E_CUDA);
  e->n_sms = prop.multiProcessorCount;
  e->smem_optin = prop.sharedMemPerBlockOptin;
  e->smem_sm = prop.sharedMemPerMultiprocessor;
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(MM_E_CUDA);
  if (cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(MM_E_CUDA);
  if (cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&e->ev_copy, cudaEventDisableTiming) != cudaSuccess)
    return bail(MM_E_CUDA);
  for (auto& ev : e->ev)
    if (cudaEventCreate(&ev) != cudaSuccess) return bail(MM_E_CUDA);
  if ((rc = build_tables(e))) return bail(rc);
  // the placement kernel keeps one slot counter (and one list head) per bin in shared memory
  if (place_smem(e, 1) > e->smem_optin) e->rank_impl = 0;
  if (place_smem(e, 0) > e->smem_optin) {
    std::snprintf(e->last_err, sizeof(e->last_err), "key domain too large for shared memory: %u bins", e->Kp);
    return bail(MM_E_ARG);
  }
  {
    // NOTE: function attributes are process-global.  Every kernel gets the device's opt-in maximum so that
    // an engine with a small key domain never lowers the limit another engine of this process relies on.
    if (allow_max_smem(e, k_colscan) != cudaSuccess)
      return bail(fail(e, cudaGetLastError(), "allow_max_smem(e, k_colscan)"));
    bool ok = allow_max_smem(e, k_hist<1024>) == cudaSuccess &&
              allow_max_smem(e, k_place<0>) == cudaSuccess;
    if (ok && e->rank_impl == 1)
      ok = allow_max_smem(e, k_place<1>) == cudaSuccess;
    ok = ok && allow_max_smem(e, k_hist<512>) == cudaSuccess;
    // The TMA-fed kernel wants >= 2 ring stages next to its per-bin state; when two such CTAs
    // (512 threads each) fit in one SM, rows = 2 x SMs so barrier phases of one overlap the other.
    if (e->rank_impl == 1 && e->Kp <= 65535u) {
      for (uint32_t st = 3; st >= 2 && !e->place2_stages; --st)
        if (2 * (place2_smem(e, st) + 1024 + 256) <= e->smem_sm) { e->place2_stages = st; e->rows_per_sm = 2; e->block = 512; }
      for (uint32_t st = kMaxStages; st >= 2 && !e->place2_stages; --st)
        if (place2_smem(e, st) + 1024 <= e->smem_optin) { e->place2_stages = st; e->rows_per_sm = 1; e->block = 1024; }
    }
    if (ok && e->place2_stages) {
      while (e->hist3_stages > 2 && ((size_t)e->rows_per_sm * (hist3_smem(e, e->hist3_stages) + 1280) > e->smem_sm ||
                                     hist3_smem(e, e->hist3_stages) + 1024 > e->smem_optin))
        --e->hist3_stages;
      ok = allow_max_smem(e, k_hist3<512>) == cudaSuccess && allow_max_smem(e, k_hist3<1024>) == cudaSuccess;
    }
    if (ok && e->place2_stages) {
      ok = allow_max_smem(e, k_place2<512>) == cudaSuccess &&
           allow_max_smem(e, k_place2<1024>) == cudaSuccess;
      e->rank_impl = 3;
    }
    if (!ok) return bail(fail(e, cudaGetLastError(), "cudaFuncSetAttribute"));
  }
  if ((rc = alloc_pool(e, e->pool[0], e->capacity)) || (rc = alloc_pool(e, e->pool[1], e->capacity))) return bail(rc);
  if (cudaMalloc(&e->d_left_bits, (((size_t)e->capacity + 3 * kRound) / 32 + 64) * 4) != cudaSuccess) return bail(MM_E_CUDA);
  if (e->use_active) {
    uint64_t want = cfg->active_capacity ? cfg->active_capacity : 2ull * cfg->capacity;
    uint64_t h = 1024;
    while (h * 3 < want * 4 + 64) h <<= 1;  // load factor <= 0.75 at active_capacity
    e->hcap = h;
    for (auto& t : e->tab) {
      if (cudaMalloc(&t.kv, h * 16) != cudaSuccess) return bail(MM_E_CUDA);
    }
    if ((rc = clear_table(e, e->tab[0]))) return bail(rc);
  }
  const size_t cap = (size_t)e->capacity + 64;
  e->max_lobbies = e->capacity / e->min_L + 1;
  auto A = [&](void** p, size_t bytes) { return cudaMalloc(p, bytes) == cudaSuccess; };
  if (!A((void**)&e->d_tot, (e->Kp + 1) * 4) || !A((void**)&e->d_outbase, (e->Kp + 1) * 4) ||
      !A((void**)&e->d_binlim, (e->Kp + 1) * 4) || !A((void**)&e->d_seg, e->n_segs * sizeof(SegInfo)) ||
      !A((void**)&e->d_members, cap * 8) || !A((void**)&e->d_src_idx, cap * 4) ||
      !A((void**)&e->d_hdr, (size_t)e->max_lobbies * sizeof(mm_lobby_hdr)) ||
      !A((void**)&e->d_emit_seq, (size_t)e->max_lobbies * 4) || !A((void**)&e->d_ctr, sizeof(TickCtr)) ||
      !A((void**)&e->d_small, 64))
    return bail(fail(e, cudaGetLastError(), "cudaMalloc"));
  if (cudaMallocHost(&e->h_ctr, sizeof(TickCtr)) != cudaSuccess || cudaMallocHost(&e->h_small, 64) != cudaSuccess)
    return bail(MM_E_CUDA);
  if ((rc = alloc_tick_scratch(e))) return bail(rc);
  if (e->rank_impl == 3 && e->block == 512 && e->rows_per_sm == 2) {
    size_t sz = std::max(hist3_smem(e, e->hist3_stages), place2_smem(e, e->place2_stages));
    sz = std::max<size_t>(sz, std::max<size_t>((size_t)kEpiScratchWords * 4, colscan_smem(e)));
    int coop = 0, nb = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device);
    if (coop && allow_max_smem(e, k_tick<512>) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tick<512>, 512, sz) == cudaSuccess &&
        (uint32_t)nb * (uint32_t)e->n_sms >= e->R) {
      e->fused_ok = 1;
      e->tick_smem = sz;
    }
    cudaGetLastError();
  }
  if (cudaMemsetAsync(e->d_tot, 0, (e->Kp + 1) * 4, e->stream) != cudaSuccess) return bail(MM_E_CUDA);
  if ((rc = set_persist(e, 1024))) return bail(rc);  // clamped to the device's persisting-L2 maximum
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) return bail(MM_E_CUDA);
  *out = e;
  return MM_OK;
}

int mm_destroy(mm_engine* e) {
  if (!e) return MM_OK;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  free_pool(e->pool[0]); free_pool(e->pool[1]); free_pool(e->snap); cudaFree(e->d_left_bits);
  for (auto& t : e->tab) cudaFree(t.kv);
  cudaFree(e->d_lut); cudaFree(e->d_grp_lut); cudaFree(e->d_mode_tsize); cudaFree(e->d_seg_bin_lo); cudaFree(e->d_seg_L);
  cudaFree(e->d_M); cudaFree(e->d_tot); cudaFree(e->d_outbase); cudaFree(e->d_binlim); cudaFree(e->d_bin_seg);
  cudaFree(e->d_bin_key); cudaFree(e->d_seg); cudaFree(e->d_members); cudaFree(e->d_src_idx); cudaFree(e->d_hdr);
  cudaFree(e->d_emit_seq); cudaFree(e->d_rescnt); cudaFree(e->d_ctr); cudaFree(e->d_small);
  cudaFree(e->d_in_id); cudaFree(e->d_hslot); cudaFree(e->d_in_rating); cudaFree(e->d_in_mode); cudaFree(e->d_code);
  cudaFree(e->d_in_ts); cudaFree(e->d_blocksum);
  if (e->h_ctr) cudaFreeHost(e->h_ctr);
  if (e->h_small) cudaFreeHost(e->h_small);
  for (auto& ev : e->ev)
    if (ev) cudaEventDestroy(ev);
  if (e->stream && e->own_stream) cudaStreamDestroy(e->stream);
  if (e->d2h_stream) { cudaStreamSynchronize(e->d2h_stream); cudaStreamDestroy(e->d2h_stream); }
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  if (e->ev_copy) cudaEventDestroy(e->ev_copy);
  cudaGetLastError();
  delete e;
  return MM_OK;
}

int mm_set_stream(mm_engine* e, void* s) {
  if (!e) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  e->stream = (cudaStream_t)s;
  e->own_stream = false;
  return set_persist(e, e->persist_bytes ? (int64_t)(e->persist_bytes >> 20) : 0);
}

int mm_set_option(mm_engine* e, const char* name, int64_t value) {
  if (!e || !name) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  if (!std::strcmp(name, "rank_impl")) {
    if (value != 0 && value != 1 && value != 3) return MM_E_ARG;
    if (value == 1 && place_smem(e, 1) > e->smem_optin) return MM_E_ARG;
    if (value == 1) {
      const int s1 = (int)place_smem(e, 1);
      CK(allow_max_smem(e, k_place<1>));
    }
    if (value == 3 && !e->place2_stages) return MM_E_ARG;
    e->rank_impl = (int)value;
    return MM_OK;
  }
  if (!std::strcmp(name, "dense")) { e->dense_ok = (int)value; return MM_OK; }
  if (!std::strcmp(name, "async_results")) {
    int rcw = wait_results(e);
    if (rcw) return rcw;
    e->async_results = value != 0;
    return MM_OK;
  }
  if (!std::strcmp(name, "max_spread")) {
    // EXTENSION (policy S1): a lobby spans at most `value` rating points; < 0 restores the reference behaviour.
    // Defined on the rating-sorted partition, so MM_ORDER_RATING only (oracle: orc_run_windowed).
    if (value >= 0 && e->cfg.order_mode != MM_ORDER_RATING) return MM_E_ARG;
    if (value > 0x7FFFFFFF) return MM_E_ARG;
    e->max_spread = value < 0 ? -1 : (int32_t)value;
    return MM_OK;
  }
  if (!std::strcmp(name, "tick_impl")) { e->tick_impl = value != 0; return MM_OK; }
  if (!std::strcmp(name, "place_debug")) {  // timing experiments only: results are NOT valid
    if (value < 0 || value > 63) return MM_E_ARG;
    e->place_debug = (int)value;
    return MM_OK;
  }
  if (!std::strcmp(name, "persist_mb")) return set_persist(e, value);
  if (!std::strcmp(name, "place2_stages")) {
    if (value < 1 || value > (int64_t)kMaxStages || place2_smem(e, (uint32_t)value) + 1024 > e->smem_optin) return MM_E_ARG;
    CK(allow_max_smem(e, k_place2<512>));
    CK(allow_max_smem(e, k_place2<1024>));
    e->place2_stages = (uint32_t)value;
    return MM_OK;
  }
  if (!std::strcmp(name, "block")) {  // threads per row CTA
    if (value != 512 && value != 1024) return MM_E_ARG;
    e->block = (int)value;
    return MM_OK;
  }
  if (!std::strcmp(name, "rows_per_sm")) {
    if (value < 1 || value > 8) return MM_E_ARG;
    CK(cudaStreamSynchronize(e->stream));
    e->rows_per_sm = (int)value;
    return alloc_tick_scratch(e);
  }
  return MM_E_ARG;
}

int mm_enqueue_device(mm_engine* e, uint32_t n, const uint64_t* id, const int32_t* rating, const uint8_t* mode,
                      const uint32_t* enq_ts, uint8_t* accepted, uint32_t* n_accepted) {
  if (!e || (n && (!id || !rating || !mode))) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  return enqueue_device_locked(e, n, id, rating, mode, enq_ts, accepted, n_accepted);
}

int mm_enqueue(mm_engine* e, uint32_t n, const uint64_t* id, const int32_t* rating, const uint8_t* mode,
               const uint32_t* enq_ts, uint8_t* accepted) {
  if (!e || (n && (!id || !rating || !mode))) return MM_E_ARG;
  if (n == 0) return MM_OK;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  int rc = enq_prepare(e, n);
  if (rc) return rc;
  // Pipelined ingest: the host columns go up in chunks on the copy stream while the previous chunk's
  // claim / dedupe / append kernels run on the engine stream — the whole device side of the ingest hides
  // behind the PCIe transfer except for the last chunk.
  const uint32_t chunk = 1u << 20;
  for (uint32_t base = 0; base < n; base += chunk) {
    const uint32_t cnt = std::min(chunk, n - base);
    CK(cudaMemcpyAsync(e->d_in_id + base, id + base, (size_t)cnt * 8, cudaMemcpyHostToDevice, e->copy_stream));
    CK(cudaMemcpyAsync(e->d_in_rating + base, rating + base, (size_t)cnt * 4, cudaMemcpyHostToDevice, e->copy_stream));
    CK(cudaMemcpyAsync(e->d_in_mode + base, mode + base, (size_t)cnt, cudaMemcpyHostToDevice, e->copy_stream));
    if (enq_ts)
      CK(cudaMemcpyAsync(e->d_in_ts + base, enq_ts + base, (size_t)cnt * 4, cudaMemcpyHostToDevice, e->copy_stream));
    CK(cudaEventRecord(e->ev_copy, e->copy_stream));
    CK(cudaStreamWaitEvent(e->stream, e->ev_copy, 0));
    if ((rc = enq_claim(e, base, cnt, e->d_in_id, e->d_in_rating, e->d_in_mode))) return rc;
    if ((rc = enq_append(e, base, cnt, e->d_in_id, e->d_in_rating, e->d_in_mode, enq_ts ? e->d_in_ts : nullptr))) return rc;
  }
  rc = enq_finish(e, n, nullptr, nullptr);
  if (rc) return rc;
  if (accepted) {
    CK(cudaMemcpyAsync(accepted, e->d_code, n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  return MM_OK;
}

int mm_remove(mm_engine* e, uint32_t n, const uint64_t* id, uint32_t* n_removed) {
  if (n_removed) *n_removed = 0;
  if (!e || (n && !id)) return MM_E_ARG;
  if (!e->use_active) return MM_E_STATE;
  if (n == 0) return MM_OK;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  int rc = ensure_enq_scratch(e, n);
  if (rc) return rc;
  const Pool& p = e->pool[e->cur];
  CK(cudaMemcpyAsync(e->d_in_id, id, (size_t)n * 8, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemsetAsync(e->d_small, 0, 16, e->stream));
  k_remove<<<(n + 255) / 256, 256, 0, e->stream>>>(n, e->d_in_id, act_view(e), p.v, p.n, e->gen, e->K, e->d_small + 2);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(e->h_small, e->d_small, 16, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  const uint32_t rem = e->h_small[2];
  e->n_active -= std::min<uint64_t>(rem, e->n_active);
  e->n_tomb += rem;
  if (n_removed) *n_removed = rem;
  return MM_OK;
}

int mm_in_queue(mm_engine* e, uint32_t n, const uint64_t* id, uint8_t* out) {
  if (!e || (n && (!id || !out))) return MM_E_ARG;
  if (!e->use_active) return MM_E_STATE;
  if (n == 0) return MM_OK;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  int rc = ensure_enq_scratch(e, n);
  if (rc) return rc;
  CK(cudaMemcpyAsync(e->d_in_id, id, (size_t)n * 8, cudaMemcpyHostToDevice, e->stream));
  k_lookup<<<(n + 255) / 256, 256, 0, e->stream>>>(n, e->d_in_id, act_view(e), e->d_code);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, e->d_code, n, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return MM_OK;
}

int mm_pool_size(mm_engine* e, uint32_t* n) {
  if (!e || !n) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  *n = e->pool[e->cur].n;
  return MM_OK;
}

int mm_active_size(mm_engine* e, uint32_t* n) {
  if (!e || !n) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  *n = (uint32_t)std::min<uint64_t>(e->n_active, 0xFFFFFFFFu);
  return MM_OK;
}

int mm_results_wait(mm_engine* e) {
  if (!e) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  return wait_results(e);
}

int mm_tick_device(mm_engine* e, uint64_t now, mm_tick_stats* stats) {
  (void)now;  // strict-parity mode has no time-expanded window
  if (!e) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  { int rcw = wait_results(e); if (rcw) return rcw; }
  const uint32_t n = e->pool[e->cur].n;
  e->last_fused = use_fused(e);
  if (e->last_fused) {
    int rc = tick_fused(e, n, false);
    if (rc) return rc;
    return tick_commit(e, n, stats);
  }
  uint32_t chunk = 0;
  int rc = tick_phase_a(e, n, &chunk);
  if (rc) return rc;
  if ((rc = tick_phase_b(e, n, chunk, false))) return rc;
  return tick_commit(e, n, stats);
}

int mm_results_device(mm_engine* e, const mm_lobby_hdr** d_lobbies, const uint64_t** d_member_ids) {
  if (!e) return MM_E_ARG;
  if (d_lobbies) *d_lobbies = e->d_hdr;
  if (d_member_ids) *d_member_ids = e->d_members;
  return MM_OK;
}

int mm_tick(mm_engine* e, uint64_t now, mm_lobby_hdr* lobbies, uint32_t lobby_cap, uint64_t* member_ids,
            uint64_t member_cap, uint32_t* emit_seq, mm_tick_stats* stats) {
  (void)now;
  if (!e) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  const uint32_t n = e->pool[e->cur].n;
  uint32_t chunk = 0;
  int rc;
  if ((rc = wait_results(e))) return rc;
  // worst-case output sizes known up front -> the fused single launch is safe
  e->last_fused = use_fused(e) && (!lobbies || (uint64_t)lobby_cap >= n / e->min_L) && (!member_ids || member_cap >= n);
  if (e->last_fused) {
    if ((rc = tick_fused(e, n, emit_seq != nullptr))) return rc;
    if ((rc = tick_commit(e, n, stats))) return rc;
    goto copy_out;
  }
  rc = tick_phase_a(e, n, &chunk);
  if (rc) return rc;
  // the counts are final after phase A: check the caller's capacities before consuming
  CK(cudaMemcpyAsync(e->h_ctr, e->d_ctr, sizeof(TickCtr), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if ((lobbies && e->h_ctr->n_lobbies > lobby_cap) || (member_ids && e->h_ctr->n_matched > member_cap)) {
    std::snprintf(e->last_err, sizeof(e->last_err), "need lobby_cap >= %u, member_cap >= %u", e->h_ctr->n_lobbies,
                  e->h_ctr->n_matched);
    CK(cudaMemsetAsync(e->d_tot, 0, (e->Kp + 1) * 4, e->stream));  // the epilogue that re-zeroes it will not run
    CK(cudaStreamSynchronize(e->stream));
    return MM_E_CAP;
  }
  if ((rc = tick_phase_b(e, n, chunk, emit_seq != nullptr))) return rc;
  if ((rc = tick_commit(e, n, stats))) return rc;
copy_out:
  const TickCtr& c = *e->h_ctr;
  // the tick is complete here (tick_commit synchronised the engine stream); with async_results the copies run on
  // their own stream and the call returns: the caller may ingest the next batch meanwhile (PCIe is full duplex)
  cudaStream_t cs = e->async_results ? e->d2h_stream : e->stream;
  if (lobbies && c.n_lobbies)
    CK(cudaMemcpyAsync(lobbies, e->d_hdr, (size_t)c.n_lobbies * sizeof(mm_lobby_hdr), cudaMemcpyDeviceToHost, cs));
  if (member_ids && c.n_matched)
    CK(cudaMemcpyAsync(member_ids, e->d_members, (size_t)c.n_matched * 8, cudaMemcpyDeviceToHost, cs));
  if (emit_seq && c.n_lobbies)
    CK(cudaMemcpyAsync(emit_seq, e->d_emit_seq, (size_t)c.n_lobbies * 4, cudaMemcpyDeviceToHost, cs));
  if (e->async_results) { e->results_pending = true; return MM_OK; }
  CK(cudaStreamSynchronize(e->stream));
  return MM_OK;
}

int mm_pool_read(mm_engine* e, uint32_t cap, uint64_t* id, int32_t* rating, uint8_t* mode, uint8_t* team_size,
                 uint32_t* enq_ts, uint32_t* n_out) {
  if (!e || !n_out) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  const Pool& p = e->pool[e->cur];
  const uint32_t n = p.n;
  std::vector<uint64_t> hid(n);
  std::vector<int32_t> hr(n);
  std::vector<uint8_t> hm(n), hs(n);
  std::vector<uint32_t> ht(n);
  if (n) {
    CK(cudaMemcpyAsync(hid.data(), p.v.id, (size_t)n * 8, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(hr.data(), p.v.rating, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(hm.data(), p.v.mode, n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(hs.data(), p.v.tsize, n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(ht.data(), p.v.ts, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  uint32_t k = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (hm[i] == MM_MODE_DEAD) continue;  // removed while queued; the next tick drops it
    if (k >= cap) return MM_E_CAP;
    if (id) id[k] = hid[i];
    if (rating) rating[k] = hr[i];
    if (mode) mode[k] = hm[i];
    if (team_size) team_size[k] = hs[i];
    if (enq_ts) enq_ts[k] = ht[i];
    ++k;
  }
  *n_out = k;
  return MM_OK;
}

int mm_snapshot(mm_engine* e) {
  if (!e) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  CK(cudaSetDevice(e->device));
  if (!e->snap.v.id) {
    int rc = alloc_pool(e, e->snap, e->capacity);
    if (rc) return rc;
  }
  const Pool& p = e->pool[e->cur];
  const size_t n = p.n;
  CK(cudaMemcpyAsync(e->snap.v.id, p.v.id, n * 8, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->snap.v.rating, p.v.rating, n * 4, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->snap.v.mode, p.v.mode, n, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->snap.v.tsize, p.v.tsize, n, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->snap.v.ts, p.v.ts, n * 4, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(e->snap.v.bin, p.v.bin, n * 2, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->snap.n = p.n;
  e->has_snap = true;
  return MM_OK;
}

int mm_restore(mm_engine* e) {
  if (!e) return MM_E_ARG;
  std::lock_guard<std::mutex> lk(e->mu);
  if (!e->has_snap) return MM_E_STATE;
  CK(cudaSetDevice(e->device));
  Pool& p = e->pool[e->cur];
  const size_t n = e->snap.n;
  CK(cudaMemcpyAsync(p.v.id, e->snap.v.id, n * 8, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(p.v.rating, e->snap.v.rating, n * 4, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(p.v.mode, e->snap.v.mode, n, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(p.v.tsize, e->snap.v.tsize, n, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(p.v.ts, e->snap.v.ts, n * 4, cudaMemcpyDeviceToDevice, e->stream));
  CK(cudaMemcpyAsync(p.v.bin, e->snap.v.bin, n * 2, cudaMemcpyDeviceToDevice, e->stream));
  p.n = e->snap.n;
  e->gen += 1;
  if (e->use_active && p.n) {
    k_restamp<<<(p.n + 255) / 256, 256, 0, e->stream>>>(p.v, p.n, act_view(e), e->gen);
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(e->stream));
  return MM_OK;
}

}  // extern "C"
