"""Rating-group sharding across the GPUs of one box.

(rating group, game-mode) partitions never interact in the reference: one queue +
exchange per group (search/worker.ex:46-66), workers per group (application.ex:26-40),
one Mnesia table per group (models/lobby_state.ex:15-29).  So the pool shards by group
with NO data-path collective: the Generic stage's routing rule
(generic/worker.ex:46-69) decides, on the host at ingest, which rank's engine receives a
player; every rank ticks independently; the job's result is the per-rank results merged
back into canonical (mode, group, emission) order.  Pure host logic, no compute.
"""
import numpy as np


def owner_of_group(g, n_groups, world):
    """Contiguous rating ranges per GPU: group g -> rank floor(g * P / G)."""
    return (np.asarray(g, np.int64) * world) // n_groups


def groups_of_rank(rank, n_groups, world):
    g = np.arange(n_groups)
    return g[owner_of_group(g, n_groups, world) == rank]


def group_of_rating(cfg, rating):
    """Vectorised generic/worker.ex:46-53 (first match in list order, else default)."""
    rating = np.asarray(rating, np.int64)
    grp = np.full(rating.shape, cfg.default_group, np.int64)
    unset = np.ones(rating.shape, bool)
    for g in range(cfg.n_groups):
        hit = unset & (rating >= cfg.group_lo[g]) & (rating <= cfg.group_hi[g])
        grp[hit] = g
        unset &= ~hit
    return grp


def route(cfg, rating, world):
    """-> owner rank of every player (-1: unroutable, no default group)."""
    grp = group_of_rating(cfg, rating)
    own = owner_of_group(np.maximum(grp, 0), cfg.n_groups, world)
    return np.where(grp < 0, -1, own)


def merge_results(cfg, per_rank):
    """per_rank[r] = (lobbies, member_ids[, emit_seq]) of rank r, each already in canonical
    (mode, group, emission) order and covering disjoint groups.  -> the whole job's
    (lobbies, member_ids[, emit_seq]) in canonical order with first_member re-based."""
    lobs = [np.asarray(x[0]) for x in per_rank]
    has_seq = all(len(x) > 2 and x[2] is not None for x in per_rank)
    parts_l, parts_m, parts_s = [], [], []
    off = 0
    for m in range(cfg.n_modes):
        for r, lob in enumerate(lobs):  # ranks own increasing group ranges
            sel = lob["mode"] == m
            if not sel.any():
                continue
            sub = lob[sel].copy()
            L = int(sub["n_members"][0])
            a = int(sub["first_member"][0])
            mem = np.asarray(per_rank[r][1])[a:a + L * len(sub)]
            sub["first_member"] = off + np.arange(len(sub), dtype=np.uint32) * L
            off += len(mem)
            parts_l.append(sub); parts_m.append(mem)
            if has_seq:
                parts_s.append(np.asarray(per_rank[r][2])[sel])
    if not parts_l:
        z = lobs[0][:0] if lobs else np.zeros(0)
        return (z, np.zeros(0, np.uint64)) + ((np.zeros(0, np.uint32),) if has_seq else ())
    out = (np.concatenate(parts_l), np.concatenate(parts_m))
    return out + ((np.concatenate(parts_s),) if has_seq else ())
