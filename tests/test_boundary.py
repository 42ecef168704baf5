"""EXTENSION beyond the reference: the cross-group boundary pass of the rating-window policy S1 —
players within W rating points of a group boundary that found no partner in their own group are matched with the
neighbouring group's band; when the two groups live on different ranks the band candidates travel point-to-point
(NCCL send/recv on GPUs).  The sharded runs (K ranks, in-process threads or a world-2/4 gloo group) must reproduce
the single-rank run lobby for lobby; the band walk itself is the windowed oracle's (orc_run_windowed)."""
import importlib
import os
import socket
import threading

import numpy as np
import pytest

from .fakes import OracleEngine

shard = importlib.import_module("microservice-matchmaking_b200.shard")
RATING = 1


MODES = (("1v1", 2, 1), ("2v2", 2, 2))


def make_cfg(pkg, G, capacity):
    return pkg.synth.make_config(n_groups=G, modes=MODES, order=RATING, capacity=max(int(capacity), 1))


def make_players(pkg, n, seed, G):
    """A sparse random pool with the neighbourhood of every group boundary cleared, plus hand-placed players around
    the boundaries: pairs / quadruples that only fit together ACROSS the boundary, and loners that must go home."""
    rng = np.random.default_rng(seed)
    ids, rating, _, _ = pkg.synth.gen_pool(seed, n)
    mode = rng.integers(0, 2, n).astype(np.uint8)
    los, his = pkg.synth.equal_width_groups(G)
    his = np.array(his[:-1])
    near = (np.abs(rating[:, None] - his[None, :]) <= 8).any(axis=1)
    ids, rating, mode = ids[~near], rating[~near], mode[~near]
    xi, xr, xm = [], [], []
    for g, hi in enumerate(his.tolist()):
        base = 10 ** 12 + 100 * g
        place = [(hi, 0), (hi + 1, 0)]                                     # 1v1: one on each side
        if g % 3 == 0:
            place += [(hi - 1, 1), (hi, 1), (hi + 1, 1), (hi + 1, 1)]      # 2v2: two + two, spread 2
        if g % 3 == 1:
            place += [(hi - 1, 0)]                                         # a third 1v1 player: one of the three goes home
        if g % 3 == 2:
            place += [(hi + 2, 1), (hi, 1)]                                # two 2v2 players: not enough, both go home
        for k, (r, m) in enumerate(place):
            xi.append(base + k); xr.append(r); xm.append(m)
    order = rng.permutation(len(ids) + len(xi))
    ids = np.concatenate([ids, np.array(xi, np.uint64)])[order]
    rating = np.concatenate([rating, np.array(xr, np.int32)])[order].astype(np.int32)
    mode = np.concatenate([mode, np.array(xm, np.uint8)])[order]
    return ids, rating, mode


def sharded_run(pkg, engine_cls, cfg, W, K, ids, rating, mode):
    """K ranks in one process (a thread per rank, blocking in-process comm).  -> canonical (regular lobbies merged,
    boundary lobbies as {(mode, group, members tuple)}), residual ids, stats"""
    owner = shard.route(cfg, rating, K)
    comm = shard.LocalComm()
    out = [None] * K

    def work(r):
        mine = owner == r
        cfg_r = make_cfg(pkg, cfg.n_groups, mine.sum())
        eng = engine_cls(cfg_r)
        eng.set_option("max_spread", W)
        assert (np.asarray(eng.enqueue(ids[mine], rating[mine], mode[mine])) == 1).all()
        lob, mem, _seq, st = eng.tick()
        bp = shard.boundary_pass(pkg, cfg, W, K, r, eng, comm, make_engine=engine_cls)
        out[r] = (lob, mem, bp, eng.pool_read()["id"], eng.in_queue(mem[:5]) if len(mem) else [])
        eng.close()

    th = [threading.Thread(target=work, args=(r,)) for r in range(K)]
    [t.start() for t in th]
    [t.join(300) for t in th]
    assert all(o is not None for o in out), "a rank died (see the traceback above)"
    mlob, mmem = shard.merge_results(cfg, [(o[0], o[1], None) for o in out])[:2]
    stats = {k: sum(o[2][k] for o in out) for k in ("sent", "received", "matched", "lobbies")}
    return out, mlob, mmem, stats, comm


def boundary_set(out):
    """{(mode, group, member tuple)} over the ranks' boundary lobbies."""
    s = set()
    for _lob, _mem, bp, _res, _ in out:
        hdrs = bp.get("hdr")
        if hdrs is None:
            continue
        for h in hdrs:
            f, n = int(h["first_member"]), int(h["n_members"])
            s.add((int(h["mode"]), int(h["group"]), tuple(int(x) for x in bp["member_ids"][f:f + n])))
    return s


def check_semantics(cfg, W, bset, rating_of, group_of):
    for m, g, members in bset:
        r = [rating_of[x] for x in members]
        assert max(r) - min(r) <= W                                     # the window holds across the boundary
        assert set(group_of[x] for x in members) <= {g, g + 1}          # only the two groups of this boundary
        assert all(cfg.group_hi[g] - W < rating_of[x] <= cfg.group_hi[g] + W for x in members)


@pytest.mark.parametrize("K", [2, 4, 8])
def test_sharded_boundary_pass_equals_single_rank_cpu(pkg, oracle, K):
    n, G, W = 4_000, 16, 2
    cfg = make_cfg(pkg, G, n + 1000)
    ids, rating, mode = make_players(pkg, n, 7, G)
    one, lob1, mem1, st1, _ = sharded_run(pkg, OracleEngine, cfg, W, 1, ids, rating, mode)
    many, lobK, memK, stK, comm = sharded_run(pkg, OracleEngine, cfg, W, K, ids, rating, mode)
    assert np.array_equal(lob1, lobK) and np.array_equal(mem1, memK)       # the local ticks: plain group sharding
    b1, bK = boundary_set(one), boundary_set(many)
    assert b1 == bK and len(b1) > 0 and st1["matched"] == stK["matched"] > 0
    assert st1["sent"] == 0 and stK["sent"] == stK["received"] > 0 and comm.bytes_sent > 0  # a real exchange happened
    rating_of = dict(zip(ids.tolist(), rating.tolist()))
    group_of = dict(zip(ids.tolist(), shard.group_of_rating(cfg, rating).tolist()))
    check_semantics(cfg, W, bK, rating_of, group_of)
    # every player is in exactly one place: a regular lobby, a boundary lobby, or still queued at home
    resid = np.concatenate([o[3] for o in many])
    taken = np.array([x for _, _, mem in bK for x in mem], np.uint64)
    allout = np.concatenate([memK, taken, resid])
    assert len(allout) == len(ids) and np.array_equal(np.sort(allout), np.sort(ids))


def test_bands_must_be_disjoint(pkg):
    cfg = pkg.synth.make_config(n_groups=32, order=RATING, capacity=10)
    with pytest.raises(ValueError):
        shard.check_bands(cfg, 100)  # groups are 157 wide
    shard.check_bands(cfg, 78)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _gloo_worker(rank, world, port, n, W, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("microservice-matchmaking_b200")
    from tests.test_boundary import make_cfg, make_players
    cfg = make_cfg(pkg, 16, n)
    ids, rating, mode = make_players(pkg, n, 7, 16)
    mine = shard.route(cfg, rating, world) == rank
    eng = OracleEngine(make_cfg(pkg, 16, mine.sum()))
    eng.set_option("max_spread", W)
    eng.enqueue(ids[mine], rating[mine], mode[mine])
    eng.tick()
    comm = shard.DistComm()
    bp = shard.boundary_pass(pkg, cfg, W, world, rank, eng, comm, make_engine=OracleEngine)
    gathered = [None] * world
    dist.all_gather_object(gathered, (bp, comm.bytes_sent))
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_boundary_pass_over_a_gloo_process_group(pkg, oracle, world):
    import torch.multiprocessing as mp
    n, W = 4_000, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, W, q)) for r in range(world)]
    [p.start() for p in procs]
    gathered = q.get(timeout=300)
    [p.join(60) for p in procs]
    cfg = make_cfg(pkg, 16, n)
    ids, rating, mode = make_players(pkg, n, 7, 16)
    one, *_ = sharded_run(pkg, OracleEngine, cfg, W, 1, ids, rating, mode)
    got = boundary_set([(None, None, g[0], None, None) for g in gathered])
    assert got == boundary_set(one) and len(got) > 0
    assert sum(g[1] for g in gathered) > 0  # bytes crossed the process group


@pytest.mark.gpu
@pytest.mark.parametrize("K", [1, 4])
def test_boundary_pass_on_gpu_engines(pkg, oracle, K):
    """Real engines (mm_take, band engines on the device) against the CPU restatement."""
    n, G, W = 9_000, 32, 2
    cfg = make_cfg(pkg, G, n)
    ids, rating, mode = make_players(pkg, n, 11, G)
    ref, lob1, mem1, st1, _ = sharded_run(pkg, OracleEngine, cfg, W, 1, ids, rating, mode)
    got, lobK, memK, stK, _ = sharded_run(pkg, pkg.Engine, cfg, W, K, ids, rating, mode)
    assert np.array_equal(lob1, lobK) and np.array_equal(mem1, memK)
    assert boundary_set(ref) == boundary_set(got) and stK["matched"] > 0
    assert sorted(np.concatenate([o[3] for o in ref])) == sorted(np.concatenate([o[3] for o in got]))
    for o in got:  # matched players stay in the active set until the lobby stage removes them
        assert all(o[4])
