/*
 * mm_oracle.c — CPU ORACLE: serialized restatement of the reference search stage.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE (see mm_oracle.h).  PARITY UNPINNED: no
 * reference golden vectors exist for this path and the reference cannot run here.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference/matchmaking/lib).  Players are reduced to their id: the search
 * stage only ever reads player["id"] (search/worker.ex:271-273,308); the rest of
 * the JSON document is carried through untouched.
 */
#define _POSIX_C_SOURCE 200809L
#include "mm_oracle.h"

#include <pthread.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------- */
/* generic/worker.ex:46-53 — find_rating_group_by_rating/1                    */
/* ------------------------------------------------------------------------- */
int orc_find_rating_group(const mm_config* cfg, double rating) {
  /* Enum.find(@groups, @default_rating_group, fn {from, to, _} ->
   *     rating >= from and rating <= to end)  — first match in list order.     */
  for (uint32_t g = 0; g < cfg->n_groups; ++g) {
    if (rating >= (double)cfg->group_lo[g] && rating <= (double)cfg->group_hi[g]) return (int)g;
  }
  return cfg->default_group;
}

/* generic/worker.ex:27 — Enum.at(@groups, Integer.floor_div(length(@groups), 2) + 1) */
int orc_default_group_index(uint32_t n_groups) {
  uint32_t idx = n_groups / 2 + 1;
  return idx < n_groups ? (int)idx : -1; /* Enum.at out of range -> nil */
}

/* game-lobby/worker.ex:37-39 — Enum.reduce(teams, 0, length(players) + acc)   */
uint32_t orc_required_slots(const uint16_t* team_counts, uint32_t n_teams) {
  uint32_t acc = 0;
  for (uint32_t t = 0; t < n_teams; ++t) acc += team_counts[t];
  return acc;
}

/* ------------------------------------------------------------------------- */
/* models/active_user.ex — a :set table keyed by player id                     */
/* ------------------------------------------------------------------------- */
#define ORC_EMPTY UINT64_MAX
typedef struct {
  uint64_t* slot;
  uint64_t mask;
} orc_active;

static inline uint64_t orc_hash(uint64_t x) {
  x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
  return x;
}
static int orc_active_init(orc_active* a, uint64_t n) {
  uint64_t cap = 16;
  while (cap < 2 * n + 2) cap <<= 1;
  a->slot = (uint64_t*)malloc(cap * sizeof(uint64_t));
  if (!a->slot) return -1;
  memset(a->slot, 0xFF, cap * sizeof(uint64_t));
  a->mask = cap - 1;
  return 0;
}
/* active_user.ex:46-55 — add_user/1 (Mnesia.write on a :set = insert or overwrite) */
static void orc_active_add(orc_active* a, uint64_t id) {
  uint64_t h = orc_hash(id) & a->mask;
  while (a->slot[h] != ORC_EMPTY && a->slot[h] != id) h = (h + 1) & a->mask;
  a->slot[h] = id;
}
/* active_user.ex:33-44 — in_queue?/1                                          */
static inline int orc_in_queue(const orc_active* a, uint64_t id) {
  uint64_t h = orc_hash(id) & a->mask;
  while (a->slot[h] != ORC_EMPTY) {
    if (a->slot[h] == id) return 1;
    h = (h + 1) & a->mask;
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* "grouped-players": %{team_name => [player, ...]}; %{} when empty            */
/* (models/lobby_state.ex:54-56)                                               */
/* ------------------------------------------------------------------------- */
typedef struct {
  uint16_t n_teams; /* 0 = %{} */
  uint16_t count[ORC_MAX_TEAMS];
  uint64_t member[ORC_MAX_LOBBY]; /* team t at [t*S, t*S + count[t]) */
} orc_teams;

/* copy a teams document: header + the L = T*S member slots actually in use (the
 * reference re-encodes the whole O(L) document per request, lobby_state.ex:97,122) */
static inline void orc_teams_copy(orc_teams* dst, const orc_teams* src, uint32_t L) {
  memcpy(dst, src, offsetof(orc_teams, member) + (size_t)L * sizeof(uint64_t));
}

/* search/worker.ex:263-265 — get_players_count/1 */
static uint32_t orc_players_count(const orc_teams* t) { return orc_required_slots(t->count, t->n_teams); }

/* ------------------------------------------------------------------------- */
/* models/lobby_state.ex — one table per rating group, rows {id, dump, mode}   */
/* ------------------------------------------------------------------------- */
typedef struct {
  uint8_t mode;
  orc_teams teams;
} orc_row;
typedef struct {
  orc_row* rows;
  uint32_t n, cap;
  char pad[48]; /* one cache line per table: workers of different groups never share one */
} orc_lobby_table;

/* lobby_state.ex:61-104 — get_state/4: select (limit 1) a row with this game_mode,
 * delete it, return its decoded dump; {:ok, %{}} when none.                       */
static void orc_get_state(orc_lobby_table* tb, uint8_t mode, uint32_t L, orc_teams* out) {
  for (uint32_t i = 0; i < tb->n; ++i) {
    if (tb->rows[i].mode == mode) {
      orc_teams_copy(out, &tb->rows[i].teams, L);
      if (i != tb->n - 1) { /* Mnesia.delete(table, id) */
        tb->rows[i].mode = tb->rows[tb->n - 1].mode;
        orc_teams_copy(&tb->rows[i].teams, &tb->rows[tb->n - 1].teams, ORC_MAX_LOBBY);
      }
      tb->n--;
      return;
    }
  }
  out->n_teams = 0; /* get_an_empty_state/0 */
  memset(out->count, 0, sizeof(out->count));
}
/* lobby_state.ex:109-131 — update_state/5: write a row under a fresh UUID        */
static int orc_update_state(orc_lobby_table* tb, uint8_t mode, uint32_t L, const orc_teams* st) {
  if (tb->n == tb->cap) {
    uint32_t nc = tb->cap ? tb->cap * 2 : 4;
    orc_row* nr = (orc_row*)realloc(tb->rows, nc * sizeof(orc_row));
    if (!nr) return -1;
    tb->rows = nr; tb->cap = nc;
  }
  tb->rows[tb->n].mode = mode;
  orc_teams_copy(&tb->rows[tb->n].teams, st, L);
  tb->n++;
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Strategist policy S0 — stand-in for the RPC at search/worker.ex:296-306.     */
/* Request {game-mode, new-player, grouped-players}; reply {added, is_filled,   */
/* grouped-players} (:308,312-313).  S0: teams "team 1".."team T" of size S from */
/* the mode table; the player joins the first team with room; filled when all   */
/* teams are full.                                                              */
/* ------------------------------------------------------------------------- */
typedef struct { int added, is_filled; } orc_reply;
static orc_reply orc_strategist_s0(const mm_config* cfg, uint8_t mode, uint64_t player, orc_teams* tm) {
  orc_reply r = {0, 0};
  const uint16_t T = cfg->modes[mode].teams, S = cfg->modes[mode].team_size;
  if (tm->n_teams == 0) { tm->n_teams = T; memset(tm->count, 0, sizeof(tm->count)); }
  for (uint16_t t = 0; t < T; ++t) {
    if (tm->count[t] < S) { tm->member[(uint32_t)t * S + tm->count[t]++] = player; r.added = 1; break; }
  }
  r.is_filled = 1;
  for (uint16_t t = 0; t < T; ++t) if (tm->count[t] < S) { r.is_filled = 0; break; }
  return r;
}

/* search/worker.ex:267-280 — remove_inactive_players/1                          */
static int orc_remove_inactive(const orc_active* act, uint16_t S, orc_teams* tm) {
  uint32_t before = orc_players_count(tm);
  for (uint16_t t = 0; t < tm->n_teams; ++t) {
    uint16_t k = 0;
    uint64_t* m = &tm->member[(uint32_t)t * S];
    for (uint16_t i = 0; i < tm->count[t]; ++i) /* Enum.filter keeps order */
      if (orc_in_queue(act, m[i])) m[k++] = m[i];
    tm->count[t] = k;
  }
  return before != orc_players_count(tm); /* is_changed */
}

/* ------------------------------------------------------------------------- */
/* The world one serialized search worker sees                                  */
/* ------------------------------------------------------------------------- */
typedef struct {
  uint8_t mode, group;
  uint16_t n_members;
  uint32_t emit_seq;     /* input index of the player whose request completed it */
  uint64_t first_member; /* offset into members */
} orc_emit;

typedef struct {
  const mm_config* cfg;
  orc_active active;
  orc_lobby_table* tables; /* [n_groups] */
  /* emitted lobbies (the messages published at search/worker.ex:315-319)        */
  orc_emit* emits; uint32_t n_emits, cap_emits;
  uint64_t* members; uint64_t n_members, cap_members;
  uint32_t n_requeued;
  int collect; /* 0 = count only (timed legs) */
  uint64_t checksum;
} orc_world;

static int orc_emit_lobby(orc_world* w, uint8_t mode, uint8_t group, const orc_teams* tm, uint32_t seq) {
  const uint16_t S = w->cfg->modes[mode].team_size;
  uint32_t cnt = orc_players_count(tm);
  if (!w->collect) {
    w->n_emits++;
    for (uint16_t t = 0; t < tm->n_teams; ++t)
      for (uint16_t i = 0; i < tm->count[t]; ++i) w->checksum += tm->member[(uint32_t)t * S + i];
    return 0;
  }
  if (w->n_emits == w->cap_emits) {
    uint32_t nc = w->cap_emits ? w->cap_emits * 2 : 1024;
    orc_emit* ne = (orc_emit*)realloc(w->emits, (size_t)nc * sizeof(orc_emit));
    if (!ne) return -1;
    w->emits = ne; w->cap_emits = nc;
  }
  if (w->n_members + cnt > w->cap_members) {
    uint64_t nc = w->cap_members ? w->cap_members * 2 : 4096;
    while (nc < w->n_members + cnt) nc *= 2;
    uint64_t* nm = (uint64_t*)realloc(w->members, nc * sizeof(uint64_t));
    if (!nm) return -1;
    w->members = nm; w->cap_members = nc;
  }
  orc_emit* e = &w->emits[w->n_emits++];
  e->mode = mode; e->group = group; e->n_members = (uint16_t)cnt; e->emit_seq = seq;
  e->first_member = w->n_members;
  for (uint16_t t = 0; t < tm->n_teams; ++t) /* team-major, join order inside a team */
    for (uint16_t i = 0; i < tm->count[t]; ++i) w->members[w->n_members++] = tm->member[(uint32_t)t * S + i];
  return 0;
}

/* search/worker.ex:291-324 — consume/5, one request.  Returns 1 when the player must
 * be requeued (:308-310), 0 otherwise, <0 on allocation failure.                   */
static int orc_consume(orc_world* w, uint8_t group, uint64_t player_id, uint8_t game_mode, uint32_t seq) {
  const uint32_t L = (uint32_t)w->cfg->modes[game_mode].teams * w->cfg->modes[game_mode].team_size;
  orc_teams grouped;                                             /* :295 */
  orc_get_state(&w->tables[group], game_mode, L, &grouped);
  orc_reply data = orc_strategist_s0(w->cfg, game_mode, player_id, &grouped); /* :296-306 */
  int requeue = orc_in_queue(&w->active, player_id) && !data.added;           /* :308 */
  int is_changed = orc_remove_inactive(&w->active, w->cfg->modes[game_mode].team_size, &grouped); /* :312 */
  if (data.is_filled && !is_changed) {                                          /* :313 */
    if (orc_emit_lobby(w, game_mode, group, &grouped, seq) < 0) return -1;      /* :314-319 */
  } else {
    if (orc_update_state(&w->tables[group], game_mode, L, &grouped) < 0) return -1; /* :320 */
  }
  return requeue;                                                                /* :323 ack */
}

static int orc_world_init(orc_world* w, const mm_config* cfg, uint32_t n, const uint64_t* id,
                          const uint8_t* alive, int collect) {
  memset(w, 0, sizeof(*w));
  w->cfg = cfg; w->collect = collect;
  if (orc_active_init(&w->active, n) < 0) return -1;
  /* middleware/worker.ex:65-70 — every queued player was add_user'ed at enqueue;
   * alive[i]==0 models a later ActiveUser.remove_user (game-lobby/worker.ex:80,96). */
  for (uint32_t i = 0; i < n; ++i) if (!alive || alive[i]) orc_active_add(&w->active, id[i]);
  w->tables = (orc_lobby_table*)calloc(cfg->n_groups ? cfg->n_groups : 1, sizeof(orc_lobby_table));
  return w->tables ? 0 : -1;
}
static void orc_world_free(orc_world* w) {
  if (w->tables) { for (uint32_t g = 0; g < w->cfg->n_groups; ++g) free(w->tables[g].rows); free(w->tables); }
  free(w->active.slot); free(w->emits); free(w->members);
}

/* ------------------------------------------------------------------------- */
/* config checks + canonical feed order                                         */
/* ------------------------------------------------------------------------- */
static int orc_check_cfg(const mm_config* cfg) {
  if (!cfg || cfg->n_groups == 0 || cfg->n_groups > MM_MAX_GROUPS) return MM_E_ARG;
  if (cfg->n_modes == 0 || cfg->n_modes > MM_MAX_MODES) return MM_E_ARG;
  if (cfg->default_group >= (int32_t)cfg->n_groups) return MM_E_ARG;
  for (uint32_t m = 0; m < cfg->n_modes; ++m) {
    uint32_t T = cfg->modes[m].teams, S = cfg->modes[m].team_size;
    if (T == 0 || S == 0 || T > ORC_MAX_TEAMS || T * S > ORC_MAX_LOBBY) return MM_E_ARG;
  }
  return MM_OK;
}

static void orc_rating_span(const mm_config* cfg, int32_t* rmin, int32_t* rmax) {
  int32_t lo = cfg->group_lo[0], hi = cfg->group_hi[0];
  for (uint32_t g = 1; g < cfg->n_groups; ++g) {
    if (cfg->group_lo[g] < lo) lo = cfg->group_lo[g];
    if (cfg->group_hi[g] > hi) hi = cfg->group_hi[g];
  }
  *rmin = lo; *rmax = hi;
}

/* feed[k] = input index of the k-th request.  ARRIVAL: identity.  RATING: stable
 * sort by (mode, clamp(rating, rmin-1, rmax+1)) — ties keep enqueue order.          */
static uint32_t* orc_feed_order(const mm_config* cfg, uint32_t order_mode, uint32_t n,
                                const int32_t* rating, const uint8_t* mode) {
  uint32_t* feed = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t));
  if (!feed) return NULL;
  if (order_mode == MM_ORDER_ARRIVAL) { for (uint32_t i = 0; i < n; ++i) feed[i] = i; return feed; }
  int32_t rmin, rmax; orc_rating_span(cfg, &rmin, &rmax);
  const uint64_t KR = (uint64_t)((int64_t)rmax - rmin + 3);
  const uint64_t K = KR * 256;
  uint32_t* cnt = (uint32_t*)calloc(K + 1, sizeof(uint32_t));
  if (!cnt) { free(feed); return NULL; }
#define ORC_KEY(i) ((uint64_t)mode[i] * KR + (uint64_t)((rating[i] < rmin - 1 ? rmin - 1 : (rating[i] > rmax + 1 ? rmax + 1 : rating[i])) - (rmin - 1)))
  for (uint32_t i = 0; i < n; ++i) cnt[ORC_KEY(i) + 1]++;
  for (uint64_t k = 0; k < K; ++k) cnt[k + 1] += cnt[k];
  for (uint32_t i = 0; i < n; ++i) feed[cnt[ORC_KEY(i)]++] = i;
#undef ORC_KEY
  free(cnt);
  return feed;
}

/* ------------------------------------------------------------------------- */
/* result assembly: canonical (mode, group, emission) order                     */
/* ------------------------------------------------------------------------- */
void orc_result_free(orc_result* r) {
  if (!r) return;
  free(r->lobbies); free(r->member_ids); free(r->emit_seq); free(r->emission_rank); free(r->residual_ids);
  memset(r, 0, sizeof(*r));
}

static int orc_assemble(const mm_config* cfg, const orc_emit* emits, uint32_t n_emits,
                        const uint64_t* members, uint64_t n_members, orc_result* out) {
  out->n_lobbies = n_emits; out->n_matched = n_members;
  out->lobbies = (mm_lobby_hdr*)malloc(((size_t)n_emits + 1) * sizeof(mm_lobby_hdr));
  out->member_ids = (uint64_t*)malloc((n_members + 1) * sizeof(uint64_t));
  out->emit_seq = (uint32_t*)malloc(((size_t)n_emits + 1) * sizeof(uint32_t));
  out->emission_rank = (uint32_t*)malloc(((size_t)n_emits + 1) * sizeof(uint32_t));
  if (!out->lobbies || !out->member_ids || !out->emit_seq || !out->emission_rank) return MM_E_CAP;
  /* stable counting sort of the emitted lobbies by (mode, group) */
  const uint32_t NS = cfg->n_modes * cfg->n_groups;
  uint32_t* start = (uint32_t*)calloc((size_t)NS + 1, sizeof(uint32_t));
  if (!start) return MM_E_CAP;
  for (uint32_t i = 0; i < n_emits; ++i) start[(uint32_t)emits[i].mode * cfg->n_groups + emits[i].group + 1]++;
  for (uint32_t s = 0; s < NS; ++s) start[s + 1] += start[s];
  uint32_t* pos = (uint32_t*)malloc(((size_t)n_emits + 1) * sizeof(uint32_t));
  if (!pos) { free(start); return MM_E_CAP; }
  for (uint32_t i = 0; i < n_emits; ++i) pos[start[(uint32_t)emits[i].mode * cfg->n_groups + emits[i].group]++] = i;
  uint64_t off = 0;
  for (uint32_t c = 0; c < n_emits; ++c) {
    const orc_emit* e = &emits[pos[c]];
    out->lobbies[c].first_member = (uint32_t)off;
    out->lobbies[c].n_members = e->n_members;
    out->lobbies[c].mode = e->mode; out->lobbies[c].group = e->group;
    out->emit_seq[c] = e->emit_seq;
    out->emission_rank[c] = pos[c];
    memcpy(&out->member_ids[off], &members[e->first_member], (size_t)e->n_members * sizeof(uint64_t));
    off += e->n_members;
  }
  free(pos); free(start);
  return MM_OK;
}

/* ------------------------------------------------------------------------- */
int orc_run_literal(const mm_config* cfg, uint32_t order_mode, uint32_t n, const uint64_t* id,
                    const int32_t* rating, const uint8_t* mode, const uint8_t* alive, orc_result* out) {
  int rc = orc_check_cfg(cfg);
  if (rc) return rc;
  memset(out, 0, sizeof(*out));
  for (uint32_t i = 0; i < n; ++i) {
    if (mode[i] >= cfg->n_modes) return MM_E_ARG;
    if (orc_find_rating_group(cfg, (double)rating[i]) < 0) return MM_E_ARG; /* MatchError, generic/worker.ex:57 */
  }
  orc_world w;
  if (orc_world_init(&w, cfg, n, id, alive, 1) < 0) return MM_E_CAP;
  uint32_t* feed = orc_feed_order(cfg, order_mode, n, rating, mode);
  if (!feed) { orc_world_free(&w); return MM_E_CAP; }
  /* requests a refusing strategist would send round the requeue loop
   * (requeue/worker.ex:41-54 -> generic -> same group queue) are re-fed after the
   * first pass; under S0 there are none.                                            */
  uint32_t n_feed = n, cap_feed = n;
  for (uint32_t k = 0; k < n_feed; ++k) {
    uint32_t i = feed[k];
    uint8_t g = (uint8_t)orc_find_rating_group(cfg, (double)rating[i]); /* generic/worker.ex:57 */
    int rq = orc_consume(&w, g, id[i], mode[i], i);
    if (rq < 0) { free(feed); orc_world_free(&w); return MM_E_CAP; }
    if (rq && w.n_requeued < n) { /* bounded: each player is requeued at most once here */
      if (n_feed == cap_feed) {
        cap_feed = cap_feed * 2 + 16;
        uint32_t* nf = (uint32_t*)realloc(feed, (size_t)cap_feed * sizeof(uint32_t));
        if (!nf) { free(feed); orc_world_free(&w); return MM_E_CAP; }
        feed = nf;
      }
      feed[n_feed++] = i; w.n_requeued++;
    }
  }
  rc = orc_assemble(cfg, w.emits, w.n_emits, w.members, w.n_members, out);
  out->n_requeued = w.n_requeued;
  /* residual = players sitting in saved partial lobbies, reported in enqueue order */
  if (rc == MM_OK) {
    uint8_t* in_partial = (uint8_t*)calloc((size_t)n + 1, 1);
    orc_active part; int ok = in_partial && orc_active_init(&part, n) == 0;
    if (!ok) { free(in_partial); free(feed); orc_world_free(&w); return MM_E_CAP; }
    uint32_t nres = 0;
    for (uint32_t g = 0; g < cfg->n_groups; ++g)
      for (uint32_t r = 0; r < w.tables[g].n; ++r) {
        const orc_teams* tm = &w.tables[g].rows[r].teams;
        const uint16_t S = cfg->modes[w.tables[g].rows[r].mode].team_size;
        for (uint16_t t = 0; t < tm->n_teams; ++t)
          for (uint16_t i2 = 0; i2 < tm->count[t]; ++i2) { orc_active_add(&part, tm->member[(uint32_t)t * S + i2]); nres++; }
      }
    out->residual_ids = (uint64_t*)malloc(((size_t)nres + 1) * sizeof(uint64_t));
    uint32_t k = 0, dead = 0;
    for (uint32_t i = 0; i < n; ++i) {
      if (alive && !alive[i]) { dead++; continue; }
      if (out->residual_ids && orc_in_queue(&part, id[i])) out->residual_ids[k++] = id[i];
    }
    out->n_residual = k; out->n_dead = dead;
    free(part.slot); free(in_partial);
  }
  free(feed); orc_world_free(&w);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* Closed form under S0: drop dead, stable partition by (mode,      */
/* group) in feed order, lobby k of a partition = ranks [kL, (k+1)L).            */
/* ------------------------------------------------------------------------- */
int orc_run_closed_form(const mm_config* cfg, uint32_t order_mode, uint32_t n, const uint64_t* id,
                        const int32_t* rating, const uint8_t* mode, const uint8_t* alive, orc_result* out) {
  int rc = orc_check_cfg(cfg);
  if (rc) return rc;
  memset(out, 0, sizeof(*out));
  const uint32_t G = cfg->n_groups, NS = cfg->n_modes * G;
  uint8_t* grp = (uint8_t*)malloc((size_t)n + 1);
  uint32_t* feed = orc_feed_order(cfg, order_mode, n, rating, mode);
  uint32_t* cnt = (uint32_t*)calloc((size_t)NS + 1, sizeof(uint32_t));
  uint32_t* part = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t));
  uint8_t* is_res = (uint8_t*)calloc((size_t)n + 1, 1);
  if (!grp || !feed || !cnt || !part || !is_res) { rc = MM_E_CAP; goto done; }
  for (uint32_t i = 0; i < n; ++i) {
    if (mode[i] >= cfg->n_modes) { rc = MM_E_ARG; goto done; }
    int g = orc_find_rating_group(cfg, (double)rating[i]);
    if (g < 0) { rc = MM_E_ARG; goto done; }
    grp[i] = (uint8_t)g;
  }
  uint32_t dead = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (alive && !alive[i]) { dead++; continue; }
    cnt[(uint32_t)mode[i] * G + grp[i] + 1]++;
  }
  for (uint32_t s = 0; s < NS; ++s) cnt[s + 1] += cnt[s];
  {
    uint32_t* cur = (uint32_t*)malloc(((size_t)NS + 1) * sizeof(uint32_t));
    if (!cur) { rc = MM_E_CAP; goto done; }
    memcpy(cur, cnt, ((size_t)NS + 1) * sizeof(uint32_t));
    for (uint32_t k = 0; k < n; ++k) {
      uint32_t i = feed[k];
      if (alive && !alive[i]) continue;
      part[cur[(uint32_t)mode[i] * G + grp[i]]++] = i;
    }
    free(cur);
  }
  uint32_t n_lob = 0; uint64_t n_mat = 0;
  for (uint32_t s = 0; s < NS; ++s) {
    uint32_t L = (uint32_t)cfg->modes[s / G].teams * cfg->modes[s / G].team_size;
    uint32_t len = cnt[s + 1] - cnt[s];
    n_lob += len / L; n_mat += (uint64_t)(len / L) * L;
  }
  out->n_lobbies = n_lob; out->n_matched = n_mat; out->n_dead = dead;
  out->lobbies = (mm_lobby_hdr*)malloc(((size_t)n_lob + 1) * sizeof(mm_lobby_hdr));
  out->member_ids = (uint64_t*)malloc((n_mat + 1) * sizeof(uint64_t));
  out->emit_seq = (uint32_t*)malloc(((size_t)n_lob + 1) * sizeof(uint32_t));
  out->emission_rank = NULL; /* not defined by the closed form */
  if (!out->lobbies || !out->member_ids || !out->emit_seq) { rc = MM_E_CAP; goto done; }
  {
    uint32_t c = 0; uint64_t off = 0;
    for (uint32_t s = 0; s < NS; ++s) {
      uint32_t L = (uint32_t)cfg->modes[s / G].teams * cfg->modes[s / G].team_size;
      uint32_t len = cnt[s + 1] - cnt[s], nl = len / L;
      for (uint32_t k = 0; k < nl; ++k) {
        out->lobbies[c].first_member = (uint32_t)off; out->lobbies[c].n_members = (uint16_t)L;
        out->lobbies[c].mode = (uint8_t)(s / G); out->lobbies[c].group = (uint8_t)(s % G);
        for (uint32_t j = 0; j < L; ++j) out->member_ids[off + j] = id[part[cnt[s] + k * L + j]];
        out->emit_seq[c] = part[cnt[s] + k * L + L - 1];
        off += L; c++;
      }
      for (uint32_t j = nl * L; j < len; ++j) is_res[part[cnt[s] + j]] = 1;
    }
  }
  {
    uint32_t nres = (uint32_t)(cnt[NS] - n_mat);
    out->residual_ids = (uint64_t*)malloc(((size_t)nres + 1) * sizeof(uint64_t));
    if (!out->residual_ids) { rc = MM_E_CAP; goto done; }
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; ++i) if (is_res[i]) out->residual_ids[k++] = id[i];
    out->n_residual = k;
  }
done:
  free(grp); free(feed); free(cnt); free(part); free(is_res);
  if (rc) orc_result_free(out);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* EXTENSION — policy S1 (rating window), see mm_oracle.h                       */
/* ------------------------------------------------------------------------- */
int orc_run_windowed(const mm_config* cfg, int32_t max_spread, uint32_t n, const uint64_t* id, const int32_t* rating,
                     const uint8_t* mode, const uint8_t* alive, orc_result* out) {
  int rc = orc_check_cfg(cfg);
  if (rc) return rc;
  memset(out, 0, sizeof(*out));
  const uint32_t G = cfg->n_groups, NS = cfg->n_modes * G;
  int32_t rmin, rmax; orc_rating_span(cfg, &rmin, &rmax);
  uint8_t* grp = (uint8_t*)malloc((size_t)n + 1);
  uint32_t* feed = orc_feed_order(cfg, MM_ORDER_RATING, n, rating, mode);
  uint32_t* cnt = (uint32_t*)calloc((size_t)NS + 1, sizeof(uint32_t));
  uint32_t* part = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t));
  uint8_t* is_res = (uint8_t*)calloc((size_t)n + 1, 1);
  orc_emit* emits = NULL; uint64_t* members = NULL;
  if (!grp || !feed || !cnt || !part || !is_res) { rc = MM_E_CAP; goto done; }
  for (uint32_t i = 0; i < n; ++i) {
    if (mode[i] >= cfg->n_modes) { rc = MM_E_ARG; goto done; }
    int g = orc_find_rating_group(cfg, (double)rating[i]);
    if (g < 0) { rc = MM_E_ARG; goto done; }
    grp[i] = (uint8_t)g;
  }
  uint32_t dead = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (alive && !alive[i]) { dead++; continue; }
    cnt[(uint32_t)mode[i] * G + grp[i] + 1]++;
  }
  for (uint32_t s = 0; s < NS; ++s) cnt[s + 1] += cnt[s];
  {
    uint32_t* cur = (uint32_t*)malloc(((size_t)NS + 1) * sizeof(uint32_t));
    if (!cur) { rc = MM_E_CAP; goto done; }
    memcpy(cur, cnt, ((size_t)NS + 1) * sizeof(uint32_t));
    for (uint32_t k = 0; k < n; ++k) {
      uint32_t i = feed[k];
      if (alive && !alive[i]) continue;
      part[cur[(uint32_t)mode[i] * G + grp[i]]++] = i;
    }
    free(cur);
  }
  emits = (orc_emit*)malloc(((size_t)n + 1) * sizeof(orc_emit));
  members = (uint64_t*)malloc(((size_t)n + 1) * sizeof(uint64_t));
  if (!emits || !members) { rc = MM_E_CAP; goto done; }
  uint32_t n_emits = 0; uint64_t n_mem = 0;
#define ORC_CK(i) ((rating[i] < rmin - 1 ? rmin - 1 : (rating[i] > rmax + 1 ? rmax + 1 : rating[i])))
  for (uint32_t s = 0; s < NS; ++s) {
    const uint32_t L = (uint32_t)cfg->modes[s / G].teams * cfg->modes[s / G].team_size;
    const uint32_t a = cnt[s], len = cnt[s + 1] - cnt[s];
    uint32_t i = 0;
    while (i + L <= len) {
      const int64_t spread = (int64_t)ORC_CK(part[a + i + L - 1]) - (int64_t)ORC_CK(part[a + i]);
      if (max_spread < 0 || spread <= max_spread) {
        orc_emit* e = &emits[n_emits++];
        e->mode = (uint8_t)(s / G); e->group = (uint8_t)(s % G); e->n_members = (uint16_t)L;
        e->emit_seq = part[a + i + L - 1]; e->first_member = n_mem;
        for (uint32_t j = 0; j < L; ++j) members[n_mem++] = id[part[a + i + j]];
        i += L;
      } else {
        is_res[part[a + i]] = 1;
        i += 1;
      }
    }
    for (; i < len; ++i) is_res[part[a + i]] = 1;
  }
#undef ORC_CK
  rc = orc_assemble(cfg, emits, n_emits, members, n_mem, out);
  if (rc == MM_OK) {
    uint32_t nres = 0;
    for (uint32_t i = 0; i < n; ++i) nres += is_res[i];
    out->residual_ids = (uint64_t*)malloc(((size_t)nres + 1) * sizeof(uint64_t));
    if (!out->residual_ids) { rc = MM_E_CAP; goto done; }
    uint32_t k = 0;
    for (uint32_t i = 0; i < n; ++i) if (is_res[i]) out->residual_ids[k++] = id[i];
    out->n_residual = k; out->n_dead = dead;
  }
done:
  free(grp); free(feed); free(cnt); free(part); free(is_res); free(emits); free(members);
  if (rc) orc_result_free(out);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* timed legs                                                                   */
/* ------------------------------------------------------------------------- */
/* Persistent session: the reference's state KEPT ACROSS REQUESTS.               */
/* LobbyState rows survive between batches (models/lobby_state.ex:61-131), the   */
/* active set is mutated by add_user / remove_user (models/active_user.ex:46-66), */
/* requests are consumed one at a time in arrival order (search/worker.ex:291-324)*/
/* — what a single serialized search worker per group does over its lifetime.    */
/* Pins the one documented deviation of the batched tick: after a  */
/* member of a SAVED partial lobby leaves, the reference fills the hole in that   */
/* team first; the tick re-derives teams from pool order.  Lobbies that saw such a */
/* hole are flagged.                                                              */
/* ------------------------------------------------------------------------- */
#define ORC_TOMB (UINT64_MAX - 1)
typedef struct { uint64_t* slot; uint64_t mask, used, live; } orc_set;
static int orc_set_init(orc_set* a, uint64_t cap) {
  uint64_t c = 16;
  while (c < cap) c <<= 1;
  a->slot = (uint64_t*)malloc(c * sizeof(uint64_t));
  if (!a->slot) return -1;
  memset(a->slot, 0xFF, c * sizeof(uint64_t));
  a->mask = c - 1; a->used = 0; a->live = 0;
  return 0;
}
static int orc_set_has(const orc_set* a, uint64_t id) {
  uint64_t h = orc_hash(id) & a->mask;
  while (a->slot[h] != ORC_EMPTY) { if (a->slot[h] == id) return 1; h = (h + 1) & a->mask; }
  return 0;
}
static int orc_set_add(orc_set* a, uint64_t id);
static int orc_set_grow(orc_set* a) {
  orc_set b;
  if (orc_set_init(&b, (a->live + 8) * 4) < 0) return -1;
  for (uint64_t i = 0; i <= a->mask; ++i) if (a->slot[i] < ORC_TOMB) orc_set_add(&b, a->slot[i]);
  free(a->slot); *a = b;
  return 0;
}
static int orc_set_add(orc_set* a, uint64_t id) { /* caller checked !has */
  if ((a->used + 1) * 2 > a->mask + 1 && orc_set_grow(a) < 0) return -1;
  uint64_t h = orc_hash(id) & a->mask;
  while (a->slot[h] != ORC_EMPTY) h = (h + 1) & a->mask;
  a->slot[h] = id; a->used++; a->live++;
  return 0;
}
static int orc_set_del(orc_set* a, uint64_t id) {
  uint64_t h = orc_hash(id) & a->mask;
  while (a->slot[h] != ORC_EMPTY) {
    if (a->slot[h] == id) { a->slot[h] = ORC_TOMB; a->live--; return 1; }
    h = (h + 1) & a->mask;
  }
  return 0;
}

typedef struct { uint8_t mode, hole; orc_teams teams; } orc_prow;
struct orc_session {
  mm_config cfg;
  orc_set active;
  orc_prow** rows; uint32_t* n_rows; uint32_t* cap_rows;   /* per group: LobbyState table */
  orc_emit* emits; uint32_t n_emits, cap_emits;
  uint8_t* holes;                                           /* per emitted lobby */
  uint64_t* members; uint64_t n_members, cap_members;
  uint32_t seq;                                             /* requests seen so far */
};

orc_session* orc_session_new(const mm_config* cfg) {
  if (orc_check_cfg(cfg) != MM_OK) return NULL;
  orc_session* s = (orc_session*)calloc(1, sizeof(orc_session));
  if (!s) return NULL;
  s->cfg = *cfg;
  s->rows = (orc_prow**)calloc(cfg->n_groups, sizeof(orc_prow*));
  s->n_rows = (uint32_t*)calloc(cfg->n_groups, sizeof(uint32_t));
  s->cap_rows = (uint32_t*)calloc(cfg->n_groups, sizeof(uint32_t));
  if (!s->rows || !s->n_rows || !s->cap_rows || orc_set_init(&s->active, 1024) < 0) { orc_session_free(s); return NULL; }
  return s;
}
void orc_session_free(orc_session* s) {
  if (!s) return;
  if (s->rows) for (uint32_t g = 0; g < s->cfg.n_groups; ++g) free(s->rows[g]);
  free(s->rows); free(s->n_rows); free(s->cap_rows); free(s->active.slot);
  free(s->emits); free(s->holes); free(s->members); free(s);
}

/* one request through consume/5 with the LobbyState of its group kept from earlier requests */
static int orc_session_consume(orc_session* s, uint8_t group, uint64_t player, uint8_t mode) {
  const mm_config* cfg = &s->cfg;
  const uint16_t S = cfg->modes[mode].team_size;
  const uint32_t L = (uint32_t)cfg->modes[mode].teams * S;
  orc_teams grouped; uint8_t hole = 0;
  grouped.n_teams = 0; memset(grouped.count, 0, sizeof(grouped.count));
  orc_prow* tb = s->rows[group];
  for (uint32_t i = 0; i < s->n_rows[group]; ++i)                    /* get_state: pop a row of this mode */
    if (tb[i].mode == mode) {
      orc_teams_copy(&grouped, &tb[i].teams, L); hole = tb[i].hole;
      tb[i] = tb[s->n_rows[group] - 1]; s->n_rows[group]--;
      break;
    }
  orc_reply data = orc_strategist_s0(cfg, mode, player, &grouped);   /* :296-306 */
  /* :308-310 requeue: never under S0 (added is always true for a non-full lobby)                  */
  /* remove_inactive_players/1 (:267-280) against the LIVE active set                              */
  uint32_t before = orc_players_count(&grouped);
  for (uint16_t t = 0; t < grouped.n_teams; ++t) {
    uint16_t k = 0; uint64_t* m = &grouped.member[(uint32_t)t * S];
    for (uint16_t i = 0; i < grouped.count[t]; ++i) if (orc_set_has(&s->active, m[i])) m[k++] = m[i];
    grouped.count[t] = k;
  }
  const int is_changed = before != orc_players_count(&grouped);
  if (is_changed) hole = 1;
  if (data.is_filled && !is_changed) {                               /* :313-319 emit */
    if (s->n_emits == s->cap_emits) {
      uint32_t nc = s->cap_emits ? s->cap_emits * 2 : 256;
      orc_emit* ne = (orc_emit*)realloc(s->emits, (size_t)nc * sizeof(orc_emit));
      uint8_t* nh = (uint8_t*)realloc(s->holes, nc);
      if (ne) s->emits = ne;
      if (nh) s->holes = nh;
      if (!ne || !nh) return -1;
      s->cap_emits = nc;
    }
    if (s->n_members + L > s->cap_members) {
      uint64_t nc = s->cap_members ? s->cap_members * 2 : 4096;
      while (nc < s->n_members + L) nc *= 2;
      uint64_t* nm = (uint64_t*)realloc(s->members, nc * sizeof(uint64_t));
      if (!nm) return -1;
      s->members = nm; s->cap_members = nc;
    }
    orc_emit* e = &s->emits[s->n_emits];
    s->holes[s->n_emits++] = hole;
    e->mode = mode; e->group = group; e->n_members = (uint16_t)orc_players_count(&grouped); e->emit_seq = s->seq;
    e->first_member = s->n_members;
    for (uint16_t t = 0; t < grouped.n_teams; ++t)
      for (uint16_t i = 0; i < grouped.count[t]; ++i) s->members[s->n_members++] = grouped.member[(uint32_t)t * S + i];
  } else {                                                            /* :320 save_new_state */
    if (s->n_rows[group] == s->cap_rows[group]) {
      uint32_t nc = s->cap_rows[group] ? s->cap_rows[group] * 2 : 4;
      orc_prow* nr = (orc_prow*)realloc(s->rows[group], (size_t)nc * sizeof(orc_prow));
      if (!nr) return -1;
      s->rows[group] = nr; s->cap_rows[group] = nc;
    }
    orc_prow* r = &s->rows[group][s->n_rows[group]++];
    r->mode = mode; r->hole = hole; orc_teams_copy(&r->teams, &grouped, L);
  }
  return 0;
}

/* middleware/worker.ex:65-70 (dedupe + add_user) -> generic/worker.ex:46-69 (routing) -> consume/5, per request */
int orc_session_feed(orc_session* s, uint32_t n, const uint64_t* id, const int32_t* rating, const uint8_t* mode,
                     uint8_t* accepted) {
  if (!s) return MM_E_ARG;
  for (uint32_t i = 0; i < n; ++i, ++s->seq) {
    const int g = mode[i] < s->cfg.n_modes ? orc_find_rating_group(&s->cfg, (double)rating[i]) : -1;
    if (g < 0) { if (accepted) accepted[i] = 2; continue; }
    if (orc_set_has(&s->active, id[i])) { if (accepted) accepted[i] = 0; continue; }  /* "You are already in the queue." */
    if (orc_set_add(&s->active, id[i]) < 0) return MM_E_CAP;
    if (accepted) accepted[i] = 1;
    if (orc_session_consume(s, (uint8_t)g, id[i], mode[i]) < 0) return MM_E_CAP;
  }
  return MM_OK;
}
/* ActiveUser.remove_user/1 (models/active_user.ex:57-66): the entry goes, saved lobbies are NOT touched */
int orc_session_remove(orc_session* s, uint32_t n, const uint64_t* id, uint32_t* n_removed) {
  if (!s) return MM_E_ARG;
  uint32_t k = 0;
  for (uint32_t i = 0; i < n; ++i) k += (uint32_t)orc_set_del(&s->active, id[i]);
  if (n_removed) *n_removed = k;
  return MM_OK;
}
/* lobbies emitted since the last take, canonical order; hole[c] = 1 when the lobby saw a mid-lobby leaver.
 * residual_ids = members of the saved partial lobbies that are still active (team-major per row; order is not
 * the enqueue order).  emit_seq = request number (all requests of the session, rejected ones included).        */
int orc_session_take(orc_session* s, orc_result* out, uint8_t** hole) {
  if (!s || !out) return MM_E_ARG;
  memset(out, 0, sizeof(*out));
  int rc = orc_assemble(&s->cfg, s->emits, s->n_emits, s->members, s->n_members, out);
  if (rc != MM_OK) return rc;
  if (hole) {
    *hole = (uint8_t*)malloc((size_t)s->n_emits + 1);
    if (!*hole) return MM_E_CAP;
    for (uint32_t c = 0; c < s->n_emits; ++c) (*hole)[c] = s->holes[out->emission_rank[c]];
  }
  uint64_t cap = 0;
  for (uint32_t g = 0; g < s->cfg.n_groups; ++g) cap += (uint64_t)s->n_rows[g] * ORC_MAX_LOBBY;
  out->residual_ids = (uint64_t*)malloc((cap + 1) * sizeof(uint64_t));
  if (!out->residual_ids) return MM_E_CAP;
  uint32_t k = 0;
  for (uint32_t g = 0; g < s->cfg.n_groups; ++g)
    for (uint32_t r = 0; r < s->n_rows[g]; ++r) {
      const orc_prow* row = &s->rows[g][r];
      const uint16_t S = s->cfg.modes[row->mode].team_size;
      for (uint16_t t = 0; t < row->teams.n_teams; ++t)
        for (uint16_t i = 0; i < row->teams.count[t]; ++i) {
          const uint64_t m = row->teams.member[(uint32_t)t * S + i];
          if (orc_set_has(&s->active, m)) out->residual_ids[k++] = m;
        }
    }
  out->n_residual = k;
  s->n_emits = 0; s->n_members = 0;
  return MM_OK;
}
void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------- */
static double orc_now(void) {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
  orc_world* w; /* shared read-only active set; private tables/emit counters */
  const uint64_t* id; const uint8_t* mode; const uint8_t* grp;
  const uint32_t* feed; /* routed: requests of this worker's groups, in feed order */
  uint32_t n_feed;
  uint32_t n_emits; uint64_t checksum;
} orc_job;

static void* orc_job_main(void* p) {
  orc_job* j = (orc_job*)p;
  orc_world lw = *j->w; /* shares active + tables (disjoint groups per thread) */
  lw.n_emits = 0; lw.checksum = 0; lw.collect = 0;
  for (uint32_t k = 0; k < j->n_feed; ++k) {
    uint32_t i = j->feed[k];
    orc_consume(&lw, j->grp[i], j->id[i], j->mode[i], i);
  }
  j->n_emits = lw.n_emits; j->checksum = lw.checksum;
  return NULL;
}

double orc_time_literal(const mm_config* cfg, uint32_t order_mode, uint32_t n, const uint64_t* id,
                        const int32_t* rating, const uint8_t* mode, uint32_t n_threads, uint32_t* n_lobbies) {
  if (orc_check_cfg(cfg) || n_threads == 0) return -1.0;
  if (n_threads > cfg->n_groups) n_threads = cfg->n_groups; /* one worker per group at most */
  orc_world w;
  if (orc_world_init(&w, cfg, n, id, NULL, 0) < 0) return -1.0;
  uint32_t* feed = orc_feed_order(cfg, order_mode, n, rating, mode);
  uint8_t* grp = (uint8_t*)malloc((size_t)n + 1);
  uint32_t* routed = (uint32_t*)malloc(((size_t)n + 1) * sizeof(uint32_t));
  orc_job* jobs = (orc_job*)calloc(n_threads, sizeof(orc_job));
  pthread_t* th = (pthread_t*)calloc(n_threads, sizeof(pthread_t));
  double secs = -1.0;
  if (!feed || !grp || !routed || !jobs || !th) goto done;
  for (uint32_t i = 0; i < n; ++i) {
    int g = orc_find_rating_group(cfg, (double)rating[i]);
    if (g < 0 || mode[i] >= cfg->n_modes) goto done;
    grp[i] = (uint8_t)g;
  }
  { /* Generic stage routing (generic/worker.ex:55-69): group g -> worker g % n_threads */
    uint32_t* start = (uint32_t*)calloc((size_t)n_threads + 1, sizeof(uint32_t));
    if (!start) goto done;
    for (uint32_t i = 0; i < n; ++i) start[grp[i] % n_threads + 1]++;
    for (uint32_t t = 0; t < n_threads; ++t) start[t + 1] += start[t];
    for (uint32_t t = 0; t < n_threads; ++t) {
      jobs[t].w = &w; jobs[t].id = id; jobs[t].mode = mode; jobs[t].grp = grp;
      jobs[t].feed = routed + start[t]; jobs[t].n_feed = start[t + 1] - start[t];
    }
    uint32_t* cur = (uint32_t*)malloc(((size_t)n_threads + 1) * sizeof(uint32_t));
    if (!cur) { free(start); goto done; }
    memcpy(cur, start, ((size_t)n_threads + 1) * sizeof(uint32_t));
    for (uint32_t k = 0; k < n; ++k) { uint32_t i = feed[k]; routed[cur[grp[i] % n_threads]++] = i; }
    free(cur); free(start);
  }
  {
    double t0 = orc_now();
    if (n_threads == 1) orc_job_main(&jobs[0]);
    else {
      for (uint32_t t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, orc_job_main, &jobs[t]);
      for (uint32_t t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    }
    secs = orc_now() - t0;
  }
  {
    uint32_t tot = 0; uint64_t cs = 0;
    for (uint32_t t = 0; t < n_threads; ++t) { tot += jobs[t].n_emits; cs += jobs[t].checksum; }
    if (n_lobbies) *n_lobbies = tot;
    if (cs == 0x5EEDFACEull) secs += 1e-12; /* keep the loop observable */
  }
done:
  free(feed); free(grp); free(routed); free(jobs); free(th);
  orc_world_free(&w);
  return secs;
}

/* ------------------------------------------------------------------------- */
/* synthetic pools, SURVEY §8(d)                                                */
/* ------------------------------------------------------------------------- */
uint64_t orc_mix64(uint64_t z) {
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
  return z;
}
void orc_gen_pool(uint64_t seed, uint64_t first, uint32_t n, uint32_t bell, uint8_t mode_const,
                  uint64_t* id, int32_t* rating, uint8_t* mode, uint32_t* enq_ts) {
  for (uint32_t k = 0; k < n; ++k) {
    uint64_t i = first + k;
    uint64_t x = orc_mix64(seed * 0x9E3779B97F4A7C15ull + i);
    if (id) id[k] = orc_mix64(((seed + 1) * 0xD1B54A32D192ED03ull) ^ i);
    if (rating) {
      if (!bell) rating[k] = (int32_t)((x >> 32) % 5001u);
      else {
        uint64_t u = (x & 0xFFFF) + ((x >> 16) & 0xFFFF) + ((x >> 32) & 0xFFFF) + ((x >> 48) & 0xFFFF);
        rating[k] = (int32_t)((u * 5000u) / (4u * 65535u)); /* floor((u1+u2+u3+u4)*5000/4), u_j in [0,1] */
      }
    }
    if (mode) mode[k] = mode_const;
    if (enq_ts) enq_ts[k] = (uint32_t)i;
  }
}
