"""N>1 path on CPU: world_size-2/4 `gloo` process groups, one engine per rank (the
oracle-backed test double stands in for the GPU engine), players routed to ranks by
rating group with no data-path collective; the merged result must equal a single
engine over the whole pool."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from .fakes import OracleEngine


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, order, n, out_q, max_spread=-1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("microservice-matchmaking_b200")
    shard = importlib.import_module("microservice-matchmaking_b200.shard")
    cfg = pkg.synth.make_config(n_groups=8, order=order, capacity=n)
    rng = np.random.default_rng(3)
    ids, rating, _, _ = pkg.synth.gen_pool(5, n, bell=True)
    mode = rng.integers(0, 2, n).astype(np.uint8)
    # the Generic stage's routing (generic/worker.ex:46-69): each rank keeps its groups' players
    mine = shard.route(cfg, rating, world) == rank
    eng = OracleEngine(cfg)
    if max_spread >= 0:
        eng.set_option("max_spread", max_spread)
    assert eng.enqueue(ids[mine], rating[mine], mode[mine]).all()
    lob, mem, seq, st = eng.tick()
    gathered = [None] * world
    dist.all_gather_object(gathered, (lob, mem, seq, eng.pool_read()["id"], int(mine.sum())))
    if rank == 0:
        assert sum(g[4] for g in gathered) == n  # every player has exactly one owner
        mlob, mmem, _ = shard.merge_results(cfg, [(g[0], g[1], g[2]) for g in gathered])
        out_q.put((mlob, mmem, np.concatenate([g[3] for g in gathered])))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,order", [(2, 0), (2, 1), (4, 1)])
def test_sharded_equals_single_engine(pkg, oracle, world, order):
    n = 30_000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, order, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    mlob, mmem, mres = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = pkg.synth.make_config(n_groups=8, order=order, capacity=n)
    rng = np.random.default_rng(3)
    ids, rating, _, _ = pkg.synth.gen_pool(5, n, bell=True)
    mode = rng.integers(0, 2, n).astype(np.uint8)
    ref = oracle.run_literal(cfg, ids, rating, mode)
    assert np.array_equal(mlob, ref.lobbies)
    assert np.array_equal(mmem, ref.member_ids)
    assert np.array_equal(np.sort(mres), np.sort(ref.residual_ids))


def test_sharded_rating_window_needs_no_exchange(pkg, oracle):
    """EXTENSION: the S1 window is defined inside a (mode, group) partition, so it never crosses a shard boundary:
    sharded ticks with max_spread merge to the single-engine result without any collective."""
    n, world, W = 20_000, 2, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1, n, q, W)) for r in range(world)]
    for p in procs:
        p.start()
    mlob, mmem, mres = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = pkg.synth.make_config(n_groups=8, order=1, capacity=n)
    rng = np.random.default_rng(3)
    ids, rating, _, _ = pkg.synth.gen_pool(5, n, bell=True)
    mode = rng.integers(0, 2, n).astype(np.uint8)
    ref = oracle.run_windowed(cfg, W, ids, rating, mode)
    assert ref.n_residual > 100 and ref.n_lobbies > 100
    assert np.array_equal(mlob, ref.lobbies)
    assert np.array_equal(mmem, ref.member_ids)
    assert np.array_equal(np.sort(mres), np.sort(ref.residual_ids))


def test_owner_mapping(pkg):
    shard = importlib.import_module("microservice-matchmaking_b200.shard")
    assert list(shard.owner_of_group(np.arange(32), 32, 8)) == [g // 4 for g in range(32)]  # 4 groups per GPU
    assert list(shard.groups_of_rank(1, 7, 2)) == [4, 5, 6]
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS)
    assert list(shard.group_of_rating(cfg, [0, 1499, 1500, 5000, 5001, -1])) == [0, 0, 1, 6, 4, 4]
    assert list(shard.route(cfg, [0, 4999], 2)) == [0, 1]
