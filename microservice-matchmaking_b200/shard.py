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


# ======================================================================================================================
# EXTENSION (BASELINE configs[3]-[4]): the boundary pass.
#
# Not reference behaviour: the reference never lets a player leave its rating group (requeue -> generic -> same
# group queue, search/worker.ex:239-248).  Under the rating-window policy S1 (mm_set_option "max_spread" = W) a
# player near a group boundary may have no partner inside its own group while a neighbour W points away sits in the
# next group — on another GPU when the boundary coincides with a shard boundary.  After every rank's local tick:
#
#   for every pair of adjacent groups (g, g+1) and every mode: the still-queued players of g with rating > hi_g - W
#   and those of g+1 with rating <= hi_g + W form ONE band; the band is matched by the same windowed walk as a
#   partition (policy S1, oracle: orc_run_windowed on the band), candidates of g first then of g+1, each in enqueue
#   order (the tie-break inside a rating); a lobby formed in the band belongs to the LOWER group g and is emitted by the
#   rank that owns g; band players that stay unmatched simply remain queued at home.
#
# Bands must be disjoint (every group at least 2 W wide), so all boundaries are independent and every rank works in
# parallel.  Data path when g and g+1 live on different ranks: the upper rank SENDS its band candidates (id, rating,
# mode columns) to the lower one, which matches the band and sends back the ids it consumed; point-to-point over the
# process group — NCCL send/recv over NVLink on GPUs, gloo in the CPU tests.  Matched players leave their home pool
# through mm_take (they stay in the active set like any matched player).
# ======================================================================================================================
def check_bands(cfg, W):
    """Adjacent groups must be contiguous ascending rating ranges, each at least 2 W wide."""
    for g in range(cfg.n_groups):
        if cfg.group_hi[g] - cfg.group_lo[g] + 1 < 2 * W:
            raise ValueError(f"rating group {g} is narrower than 2 x max_spread: boundary bands would overlap")
        if g and cfg.group_lo[g] != cfg.group_hi[g - 1] + 1:
            raise ValueError("the boundary pass needs contiguous ascending rating groups")


def band_config(pkg, cfg, g, W, capacity):
    """One-group engine configuration covering the band around the boundary of groups g | g+1."""
    hi = int(cfg.group_hi[g])
    modes = [(f"m{m}", int(cfg.modes[m].teams), int(cfg.modes[m].team_size)) for m in range(cfg.n_modes)]
    c = pkg.synth.make_config(groups=[(hi - W + 1, hi + W)], modes=modes, order=pkg.abi.MM_ORDER_RATING,
                              capacity=max(int(capacity), 1), default_group=-1, device=int(cfg.device))
    return c


class LocalComm:
    """In-process stand-in for the point-to-point exchange (GPU tests where K engines share one process, one thread
    per rank): a blocking queue per (src, dst)."""

    def __init__(self):
        import collections
        import queue
        self.box = collections.defaultdict(queue.Queue)
        self.bytes_sent = 0

    def send(self, src, dst, arrays):
        arrays = [np.array(a, copy=True) for a in arrays]
        self.bytes_sent += sum(a.nbytes for a in arrays)
        self.box[(src, dst)].put(arrays)

    def recv(self, src, dst):
        return self.box[(src, dst)].get(timeout=120)


class DistComm:
    """torch.distributed point-to-point: a count, then the columns as tensors (on the GPU for NCCL)."""
    DTYPES = ("int64", "int32", "uint8")

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, device
        self.bytes_sent = 0

    def send(self, src, dst, arrays):
        t, d = self.torch, self.dist
        n = len(arrays[0])
        d.send(t.tensor([n, len(arrays)], dtype=t.int64, device=self.device), dst)
        for a, dt in zip(arrays, self.DTYPES):
            if n:
                x = t.from_numpy(np.ascontiguousarray(a).view(getattr(np, dt) if dt != "int64" else np.int64))
                d.send(x.to(self.device) if self.device is not None else x, dst)
                self.bytes_sent += x.numel() * x.element_size()

    def recv(self, src, dst):
        t, d = self.torch, self.dist
        hdr = t.zeros(2, dtype=t.int64, device=self.device)
        d.recv(hdr, src)
        n, k = int(hdr[0].item()), int(hdr[1].item())
        out = []
        for dt in self.DTYPES[:k]:
            x = t.zeros(n, dtype=getattr(t, dt), device=self.device)
            if n:
                d.recv(x, src)
            out.append(x.cpu().numpy())
        if out:
            out[0] = out[0].view(np.uint64)
        return out


def boundary_pass(pkg, cfg, W, world, rank, engine, comm, make_engine=None, cache=None):
    """Run this rank's share of the boundary pass (see above).  `engine` has just ticked.
    cache: dict kept by the caller across ticks — the band engines are reused instead of re-created.
    -> dict(hdr, member_ids, sent, received, matched, lobbies) — hdr: mm_lobby_hdr records (first_member into member_ids)
    carrying the LOWER group's index; sent / received / matched: players; lobbies: count."""
    check_bands(cfg, W)
    make_engine = make_engine or pkg.Engine
    G = cfg.n_groups
    mine = groups_of_rank(rank, G, world)
    res = engine.pool_read()
    ids, rating, mode = res["id"], res["rating"], res["mode"]
    grp = group_of_rating(cfg, rating)
    stats = {"sent": 0, "received": 0, "matched": 0, "lobbies": 0}
    if len(mine) == 0:
        return dict(hdr=None, member_ids=np.zeros(0, np.uint64), **stats)
    a, b = int(mine[0]), int(mine[-1])
    if rank > 0:  # my lowest group's low-side band goes down to the owner of group a-1
        sel = (grp == a) & (rating <= cfg.group_hi[a - 1] + W)
        comm.send(rank, rank - 1, (ids[sel], rating[sel], mode[sel]))
        stats["sent"] += int(sel.sum())
    lob_parts, mem_parts, off = [], [], 0
    for g in range(a, min(b, G - 2) + 1):
        lo_sel = (grp == g) & (rating > cfg.group_hi[g] - W)
        low = (ids[lo_sel], rating[lo_sel], mode[lo_sel])
        remote = g + 1 > b
        if remote:
            high = comm.recv(rank + 1, rank)
            stats["received"] += len(high[0])
        else:
            hi_sel = (grp == g + 1) & (rating <= cfg.group_hi[g] + W)
            high = (ids[hi_sel], rating[hi_sel], mode[hi_sel])
        n_band = len(low[0]) + len(high[0])
        took_remote = np.zeros(0, np.uint64)
        if n_band:
            band = cache.get(g) if cache is not None else None
            if band is None or band.cfg.capacity < n_band:
                if band is not None:
                    band.close()
                band = make_engine(band_config(pkg, cfg, g, W, max(n_band, 1 << 16 if cache is not None else 1)))
                band.set_option("max_spread", W)
                if cache is not None:
                    cache[g] = band
            for part in (low, high):  # the lower group's candidates first: tie-break inside a rating
                if len(part[0]):
                    acc = band.enqueue(part[0], part[1], part[2])
                    assert (np.asarray(acc) == 1).all()
            lob, mem, _seq, st = band.tick()
            if cache is None:
                band.close()
            else:  # leave the band engine empty for the next tick: the unmatched candidates went back home
                band.remove(np.concatenate([low[0], high[0]]))
            if len(lob):
                lob = lob.copy()
                lob["group"] = g
                lob["first_member"] = lob["first_member"] + off
                off += len(mem)
                lob_parts.append(lob); mem_parts.append(mem)
                stats["matched"] += len(mem); stats["lobbies"] += len(lob)
                own = np.isin(mem, ids)
                if own.any():
                    assert engine.take(mem[own]) == int(own.sum())
                took_remote = mem[~own]
        if remote:
            comm.send(rank, rank + 1, (took_remote,))
    if rank > 0:  # what the lower rank consumed of my candidates leaves my pool
        (gone,) = comm.recv(rank - 1, rank)
        if len(gone):
            assert engine.take(gone) == len(gone)
    hdr = np.concatenate(lob_parts) if lob_parts else None
    members = np.concatenate(mem_parts) if mem_parts else np.zeros(0, np.uint64)
    return dict(hdr=hdr, member_ids=members, **stats)
