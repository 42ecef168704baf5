/*
 * mm_oracle.h — CPU ORACLE for the search-stage hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the
 * engine (libmm_engine.so) never links, imports or falls back to it.
 *
 * PARITY UNPINNED: the reference (Elixir, /root/reference) cannot be built or run
 * in this environment (no BEAM, no RabbitMQ), the commit does not compile as
 * written (game-lobby/worker.ex:56) and the match decision itself lives in an
 * external service, openmatchmaking/microservice-strategist:0.2.3
 * (docker-compose.dev.yml:58; call site search/worker.ex:296-306), whose source is
 * not in the tree.  The reference ships no golden vectors for this path (its only
 * test is the HTTP health check, test/health_check_test.exs).  This oracle is a
 * literal, serialized restatement of the reference's own control flow with the
 * documented stand-in strategist policy S0; its known-answer tests are
 * the rules that ARE pinned by the reference's code.
 */
#ifndef MM_ORACLE_H
#define MM_ORACLE_H

#include <stdint.h>
#include "../include/mm_engine.h" /* mm_config / mm_lobby_hdr PODs only */

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_TEAMS 16
#define ORC_MAX_LOBBY 256

/* generic/worker.ex:46-53 with Erlang number semantics (rating may be a float:
 * 1499.5 falls between the integer ranges and takes the default group).           */
int orc_find_rating_group(const mm_config* cfg, double rating);

/* generic/worker.ex:27 — Enum.at(groups, div(length, 2) + 1); -1 when out of range
 * (nil in the reference, which then crashes on the MatchError at :57).            */
int orc_default_group_index(uint32_t n_groups);

/* game-lobby/worker.ex:37-39 — calculate_required_slots/1                          */
uint32_t orc_required_slots(const uint16_t* team_counts, uint32_t n_teams);

typedef struct orc_result {
  uint32_t n_lobbies;
  uint64_t n_matched;
  uint32_t n_residual;
  uint32_t n_dead;
  uint32_t n_requeued; /* players the strategist refused (never under S0) */
  /* lobbies in canonical order: (mode, group, emission order in the partition)     */
  mm_lobby_hdr* lobbies;
  uint64_t* member_ids;
  uint32_t* emit_seq;      /* per canonical lobby: input index of the completing member */
  uint32_t* emission_rank; /* per canonical lobby: its rank in global emission order    */
  uint64_t* residual_ids;  /* still queued, in enqueue order                            */
} orc_result;

void orc_result_free(orc_result* r);

/* The literal loop: feed the n queued players ONE AT A TIME, in order_mode's
 * canonical order, through a restatement of Search.Worker.consume/5
 * (search/worker.ex:291-324) against in-memory LobbyState
 * (models/lobby_state.ex) and ActiveUser (models/active_user.ex) tables.
 * alive[i]==0 marks a player removed from the active set while still queued
 * (alive may be NULL = all alive).  Returns 0 or a negative mm_status.              */
int orc_run_literal(const mm_config* cfg, uint32_t order_mode, uint32_t n,
                    const uint64_t* id, const int32_t* rating, const uint8_t* mode,
                    const uint8_t* alive, orc_result* out);

/* Same result through the closed form under S0: drop dead players, stable
 * partition by (mode, group) in feed order, cut each partition into lobbies of L. */
int orc_run_closed_form(const mm_config* cfg, uint32_t order_mode, uint32_t n,
                        const uint64_t* id, const int32_t* rating, const uint8_t* mode,
                        const uint8_t* alive, orc_result* out);

/* EXTENSION (not reference behaviour): strategist policy S1 — "a lobby may not span more
 * than max_spread rating points".  Defined directly on the sorted partition, RATING order only: per (mode, group)
 * partition sorted by (clamp(rating), enqueue order), i = 0; while i + L <= n: if key[i+L-1] - key[i] <= max_spread
 * emit players i .. i+L-1 as a lobby and i += L, else player i stays queued and i += 1.  max_spread < 0 = unlimited
 * (identical to S0).  Everything else (dead players, canonical lobby order, residual in enqueue order) as above.    */
int orc_run_windowed(const mm_config* cfg, int32_t max_spread, uint32_t n, const uint64_t* id, const int32_t* rating,
                     const uint8_t* mode, const uint8_t* alive, orc_result* out);

/* Persistent session (ARRIVAL order): LobbyState rows and the active set are KEPT across calls, requests are
 * consumed one at a time — the reference's own lifetime behaviour (models/lobby_state.ex:61-131,
 * search/worker.ex:312-321).  orc_session_take returns the lobbies emitted since the last take in canonical order
 * with hole[c] = 1 for a lobby that saw a member leave while it was being filled (free hole with orc_free).       */
typedef struct orc_session orc_session;
orc_session* orc_session_new(const mm_config* cfg);
void orc_session_free(orc_session* s);
int orc_session_feed(orc_session* s, uint32_t n, const uint64_t* id, const int32_t* rating, const uint8_t* mode,
                     uint8_t* accepted);
int orc_session_remove(orc_session* s, uint32_t n, const uint64_t* id, uint32_t* n_removed);
int orc_session_take(orc_session* s, orc_result* out, uint8_t** hole);
void orc_free(void* p);

/* Timed legs for bench.py.  Both run the LITERAL loop and return wall seconds
 * (setup — building the active set, routing by group — is outside the timer, as the
 * middleware / generic stages are outside the search stage).  n_threads==1: one
 * serialized loop.  n_threads>1: one worker per rating group as in
 * application.ex:26-40, groups dealt round-robin to threads.
 * Returns lobbies emitted in *n_lobbies.                                           */
double orc_time_literal(const mm_config* cfg, uint32_t order_mode, uint32_t n,
                        const uint64_t* id, const int32_t* rating, const uint8_t* mode,
                        uint32_t n_threads, uint32_t* n_lobbies);

/* Synthetic pool generator of SURVEY §8(d) (C twin of synth.py; tests compare).    */
uint64_t orc_mix64(uint64_t z);
void orc_gen_pool(uint64_t seed, uint64_t first, uint32_t n, uint32_t bell,
                  uint8_t mode_const, uint64_t* id, int32_t* rating, uint8_t* mode,
                  uint32_t* enq_ts);

#ifdef __cplusplus
}
#endif
#endif
