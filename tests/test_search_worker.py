"""Host mirror of Matchmaking.Search.Worker (search_worker.py): queue/exchange names,
ack-after-resident semantics, lobby JSON shape, duplicates, leavers.  The same scenario
runs on the CPU (OracleEngine test double) and on the GPU (real Engine via the C ABI)."""
import importlib
import json

import numpy as np
import pytest

from .fakes import FakeBroker, OracleEngine

sw = importlib.import_module("microservice-matchmaking_b200.search_worker")


def boot(pkg, engine_cls, order=0):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=order, capacity=10_000)
    eng = engine_cls(cfg)
    broker = FakeBroker()
    pool = sw.SearchPool(eng, ["1v1", "5v5"], pkg.synth.REFERENCE_GROUP_NAMES, flush_every_s=3600)
    workers = {}
    for g in pkg.synth.REFERENCE_GROUP_NAMES:  # application.ex:26-40: one worker per rating group
        ok, w = sw.SearchWorker.start_link(broker, pool, {"group_name": g, "channel_name": f"search.{g}"})
        assert ok == "ok"
        workers[g] = w
    broker.bind(sw.EXCHANGE_FORWARD, sw.QUEUE_FORWARD, sw.QUEUE_FORWARD)  # the lobby stage's queue
    return cfg, eng, broker, pool, workers


def publish_player(pkg, broker, cfg, pid, rating, mode, extra=None):
    """What Generic.Worker.consume/4 does (generic/worker.ex:55-69): route by rating group."""
    from oracle import oracle as orc
    gi = orc.find_rating_group(cfg, rating)
    name = pkg.synth.REFERENCE_GROUP_NAMES[gi]
    doc = {"id": pid, "rating": rating, "game-mode": mode, "response-queue": f"resp.{pid}", "event-name": "find-game"}
    doc.update(extra or {})
    broker.publish(sw.generate_exchange_name(name), sw.generate_queue_name(name), json.dumps(doc))


def scenario(pkg, engine_cls):
    cfg, eng, broker, pool, workers = boot(pkg, engine_cls)
    assert workers["gold"].config["queue"]["name"] == "matchmaking.queues.gold"  # worker.ex:46-66
    assert workers["gold"].config["exchange"]["name"] == "open-matchmaking.matchmaking.gold.direct"
    assert workers["gold"].config["qos"] == {"prefetch_count": pool.max_batch}  # not the reference's 10: see prepare_config
    players = [("u1", 100, "1v1"), ("u2", 2100, "1v1"), ("u3", 150, "1v1"), ("u4", 2200, "1v1"),
               ("u5", 4500, "1v1"), ("u1", 100, "1v1")]  # u1 twice: "already in the queue"
    for pid, r, m in players:
        publish_player(pkg, broker, cfg, pid, r, m)
    assert broker.deliver_all() == 6
    assert not broker.acked  # nothing is acked before the players are resident
    assert pool.tick() == 2
    assert len(broker.acked) == 6 and not broker.nacked
    assert pool.stats == {"enqueued": 5, "duplicates": 1, "invalid": 0, "lobbies": 2, "failed_batches": 0}
    lobbies = [json.loads(p) for p, _ in broker.queues[sw.QUEUE_FORWARD]]
    props = [pr for _, pr in broker.queues[sw.QUEUE_FORWARD]]
    assert all(pr == {"persistent": True, "content_type": "application/json"} for pr in props)  # worker.ex:254-258
    assert [set(l) for l in lobbies] == [{"teams", "game-mode"}] * 2  # worker.ex:315-318
    got = sorted(tuple(p["id"] for t in sorted(l["teams"]) for p in l["teams"][t]) for l in lobbies)
    assert got == [("u1", "u3"), ("u2", "u4")]
    l0 = lobbies[0]
    assert l0["game-mode"] == "1v1" and set(l0["teams"]) == {"team 1", "team 2"}
    p = l0["teams"]["team 1"][0]
    assert "game-mode" not in p and p["response-queue"].startswith("resp.") and p["event-name"] == "find-game"
    # the lobby stage's slot count (game-lobby/worker.ex:37-39)
    assert sum(len(v) for v in l0["teams"].values()) == 2
    # u5 is still searching; matched players stay "in queue" until the lobby stage removes them
    assert pool.in_queue("u5") and pool.in_queue("u1")
    assert pool.remove_user("u1") == ("ok", "removed") and not pool.in_queue("u1")
    ok, st = workers["grandmaster"].status()
    assert ok == "ok" and st["queue"] == "matchmaking.queues.grandmaster" and st["consumer_count"] == 1
    # a leaver is never matched (search/worker.ex:267-280)
    pool.remove_user("u5")
    publish_player(pkg, broker, cfg, "u6", 4400, "1v1")
    publish_player(pkg, broker, cfg, "u7", 4600, "1v1")
    broker.deliver_all()
    assert pool.tick() == 1
    last = json.loads(broker.queues[sw.QUEUE_FORWARD][-1][0])
    assert [p["id"] for t in ("team 1", "team 2") for p in last["teams"][t]] == ["u6", "u7"]
    # 5v5: ten players of one group -> one lobby, team 1 = first five joiners
    for i in range(10):
        publish_player(pkg, broker, cfg, f"v{i}", 3000 + i, "5v5")
    publish_player(pkg, broker, cfg, "bad", 3000, "7v7")  # unknown mode -> nack
    broker.deliver_all()
    assert pool.tick() == 1 and len(broker.nacked) == 1
    last = json.loads(broker.queues[sw.QUEUE_FORWARD][-1][0])
    assert last["game-mode"] == "5v5"
    assert [p["id"] for p in last["teams"]["team 1"]] == [f"v{i}" for i in range(5)]
    assert [p["id"] for p in last["teams"]["team 2"]] == [f"v{i}" for i in range(5, 10)]
    eng.close()


def test_worker_scenario_cpu(pkg):
    scenario(pkg, OracleEngine)


@pytest.mark.gpu
def test_worker_scenario_gpu(pkg):
    scenario(pkg, pkg.Engine)


def window_scenario(pkg, engine_cls):
    """EXTENSION: the pool-level time-expanded window (WindowSchedule) drives policy S1 tick by tick."""
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=1, capacity=10_000)
    eng = engine_cls(cfg)
    broker = FakeBroker()
    now = [100.0]
    pool = sw.SearchPool(eng, ["1v1", "5v5"], pkg.synth.REFERENCE_GROUP_NAMES,
                         window=sw.WindowSchedule(w0=10, growth_per_s=20, w_max=400), clock=lambda: now[0],
                         flush_every_s=3600)
    for g in pkg.synth.REFERENCE_GROUP_NAMES:
        sw.SearchWorker.start_link(broker, pool, {"group_name": g, "channel_name": f"search.{g}"})
    broker.bind(sw.EXCHANGE_FORWARD, sw.QUEUE_FORWARD, sw.QUEUE_FORWARD)
    for pid, r in (("a", 1000), ("b", 1008), ("c", 1100), ("d", 1300)):
        publish_player(pkg, broker, cfg, pid, r, "1v1")
    broker.deliver_all()
    assert pool.tick() == 1 and pool.last_spread == 10         # a-b are 8 apart; c, d wait
    last = json.loads(broker.queues[sw.QUEUE_FORWARD][-1][0])
    assert [p["id"] for t in ("team 1", "team 2") for p in last["teams"][t]] == ["a", "b"]
    now[0] += 2.0                                              # oldest queued (c) has waited 2 s -> W = 50
    assert pool.tick() == 0 and pool.last_spread == 50
    now[0] += 8.0                                              # 10 s -> W = 210 >= 200
    assert pool.tick() == 1 and pool.last_spread == 210
    last = json.loads(broker.queues[sw.QUEUE_FORWARD][-1][0])
    assert [p["id"] for t in ("team 1", "team 2") for p in last["teams"][t]] == ["c", "d"]
    publish_player(pkg, broker, cfg, "e", 1400, "1v1")
    broker.deliver_all()
    now[0] += 100.0
    assert pool.tick() == 0 and pool.last_spread == 10         # e just arrived: the window is tight again
    assert pool.in_queue("e") and not pool.enqueued_at.keys() - {pool.handles.lookup("e")}
    pool.remove_user("e")
    assert not pool.enqueued_at
    assert pool.tick() == 0 and pool.last_spread == 10
    eng.close()


def test_window_schedule_cpu(pkg):
    window_scenario(pkg, OracleEngine)


@pytest.mark.gpu
def test_window_schedule_gpu(pkg):
    window_scenario(pkg, pkg.Engine)


def test_start_link_contract(pkg):
    with pytest.raises(RuntimeError, match="group_name"):
        sw.prepare_config({})  # worker.ex:55-57
    pool = sw.SearchPool(OracleEngine(pkg.synth.make_config(n_groups=7)), ["1v1"], list("abcdefg"))
    assert sw.SearchWorker.start_link(None, pool, {"group_name": "gold"}) == ("error", "noconn")  # worker.ex:225-228


def test_handle_info_clauses(pkg):
    cfg, eng, broker, pool, workers = boot(pkg, OracleEngine)
    w = workers["bronze"]
    assert w.handle_info(("basic_consume_ok", {}))[0] == "noreply"
    assert w.handle_info(("basic_cancel", {}))[:2] == ("stop", "normal")
    assert w.handle_info(("basic_cancel_ok", {}))[0] == "noreply"
    assert w.handle_info(("DOWN", None))[0] == "noreply" and "consumer" in w.meta


def test_handle_table_is_dense_collision_free_and_recycles():
    t = sw.HandleTable(capacity=3)
    a, new_a = t.acquire("c0a8012e-1c9b-4b7e-9d2f-5f1d3a2b4c6d")
    assert (a, new_a) == (0, True) and t.acquire("c0a8012e-1c9b-4b7e-9d2f-5f1d3a2b4c6d") == (0, False)
    assert t.acquire("b") == (1, True) and t.acquire("c") == (2, True)
    assert t.acquire("d") == (None, False)                  # handle range exhausted
    assert t.release("b") == 1 and t.lookup("b") is None
    assert t.acquire("d") == (1, True) and len(t) == 3      # the freed handle is reused
    big = sw.HandleTable()
    assert len({big.acquire(f"p{i}")[0] for i in range(10000)}) == 10000


def test_prefetch_is_enforced_and_ingest_never_stalls(pkg):
    """with the reference's QoS (prefetch 10) a worker that acks only after the batched ingest gets 10
    deliveries per flush; with prefetch = the pool's batch size (what start_link asks for) one round delivers all.
    Either way nothing deadlocks: tick() flushes, acks, and the broker delivers the next window."""
    for prefetch, rounds_expected in ((10, lambda r: r > 10), (None, lambda r: r == 1)):
        cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=0, capacity=10_000)
        broker = FakeBroker()
        pool = sw.SearchPool(OracleEngine(cfg), ["1v1", "5v5"], pkg.synth.REFERENCE_GROUP_NAMES, flush_every_s=3600)
        for g in pkg.synth.REFERENCE_GROUP_NAMES:
            opts = {"group_name": g, "channel_name": f"search.{g}"}
            if prefetch:
                opts["prefetch_count"] = prefetch
            sw.SearchWorker.start_link(broker, pool, opts)
        for i in range(600):
            publish_player(pkg, broker, cfg, f"p{i}", 2100 + (i % 300), "1v1")  # all in "gold": one worker
        rounds = 0
        while broker.queues[sw.generate_queue_name("gold")] or pool._staged:
            delivered = broker.deliver_all()
            assert delivered <= (prefetch or pool.max_batch)
            pool.tick()
            rounds += 1
            assert rounds < 200
        assert rounds_expected(rounds) and pool.stats["enqueued"] == 600 and len(broker.acked) == 600


def test_time_triggered_flush(pkg):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=0, capacity=100)
    broker, now = FakeBroker(), [0.0]
    pool = sw.SearchPool(OracleEngine(cfg), ["1v1"], pkg.synth.REFERENCE_GROUP_NAMES, flush_every_s=0.005, clock=lambda: now[0])
    for g in pkg.synth.REFERENCE_GROUP_NAMES:
        sw.SearchWorker.start_link(broker, pool, {"group_name": g, "channel_name": f"search.{g}"})
    publish_player(pkg, broker, cfg, "a", 100, "1v1")
    broker.deliver_all()
    assert not broker.acked                     # staged, not resident yet
    now[0] += 0.010
    publish_player(pkg, broker, cfg, "b", 120, "1v1")
    broker.deliver_all()                        # the oldest staged delivery waited 10 ms: ingest now, no tick needed
    assert len(broker.acked) == 2 and pool.in_queue("a") and pool.in_queue("b")


def test_failing_engine_nacks_the_whole_batch(pkg):
    """A batch mm_enqueue refuses (active set full / CUDA error) must not vanish: every staged delivery is nacked,
    the handles taken for it are released, and the worker keeps serving."""
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=0, capacity=100)

    class Failing(OracleEngine):
        fail = True

        def enqueue(self, *a, **k):
            if self.fail:
                raise RuntimeError("mm_enqueue: status -3 — capacity exceeded")
            return super().enqueue(*a, **k)

    broker, eng = FakeBroker(), Failing(cfg)
    pool = sw.SearchPool(eng, ["1v1"], pkg.synth.REFERENCE_GROUP_NAMES, flush_every_s=3600)
    for g in pkg.synth.REFERENCE_GROUP_NAMES:
        sw.SearchWorker.start_link(broker, pool, {"group_name": g, "channel_name": f"search.{g}"})
    for i in range(5):
        publish_player(pkg, broker, cfg, f"p{i}", 100 + i, "1v1")
    broker.deliver_all()
    assert pool.flush() == 0 and len(broker.nacked) == 5 and not broker.acked
    assert pool.stats["failed_batches"] == 1 and len(pool.handles) == 0 and not pool._staged
    eng.fail = False
    publish_player(pkg, broker, cfg, "p9", 100, "1v1")
    broker.deliver_all()
    assert pool.flush() == 1 and len(broker.acked) == 1 and pool.in_queue("p9")


def test_hostile_payloads_are_nacked_not_raised(pkg):
    """json.loads accepts NaN / Infinity and non-object documents; a list game-mode is unhashable."""
    cfg, eng, broker, pool, workers = boot(pkg, OracleEngine)
    q, ex = sw.generate_queue_name("gold"), sw.generate_exchange_name("gold")
    for doc in ('[1, 2]', '"str"', '{"id": "n1", "rating": NaN, "game-mode": "1v1"}',
                '{"id": "n2", "rating": Infinity, "game-mode": "1v1"}', '{"id": "n3", "rating": 2100, "game-mode": ["1v1"]}',
                '{"id": "n4", "rating": true, "game-mode": "1v1"}', '{"id": {"x": 1}, "rating": 2100, "game-mode": "1v1"}',
                'not json at all', '{"id": "n5", "rating": 1e400, "game-mode": "1v1"}'):
        broker.publish(ex, q, doc)
    assert broker.deliver_all() == 9
    pool.flush()
    assert len(broker.nacked) == 9 and not broker.acked and pool.stats["enqueued"] == 0
    broker.publish(ex, q, '{"id": "ok", "rating": 1e3, "game-mode": "1v1"}')  # a float that is an integer is fine
    broker.deliver_all(); pool.flush()
    assert len(broker.acked) == 1


def test_malformed_and_overflowing_deliveries(pkg):
    """nack: no id / no game mode / no rating / pool full; float ratings between the integer ranges take the default
    group (generic/worker.ex:46-53); batches are flushed at max_batch without waiting for the tick."""
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=0, capacity=4)
    eng = OracleEngine(cfg)
    broker = FakeBroker()
    pool = sw.SearchPool(eng, ["1v1", "5v5"], pkg.synth.REFERENCE_GROUP_NAMES, max_batch=2)
    for g in pkg.synth.REFERENCE_GROUP_NAMES:
        sw.SearchWorker.start_link(broker, pool, {"group_name": g, "channel_name": f"search.{g}"})
    q = sw.generate_queue_name("gold")
    ex = sw.generate_exchange_name("gold")
    broker.publish(ex, q, json.dumps({"rating": 2100, "game-mode": "1v1"}))                 # no id
    broker.publish(ex, q, json.dumps({"id": "x1", "rating": 2100}))                        # no game mode
    broker.publish(ex, q, json.dumps({"id": "x2", "game-mode": "1v1"}))                    # no rating
    assert broker.deliver_all() == 3 and len(broker.nacked) == 3 and not broker.acked
    broker.publish(ex, q, json.dumps({"id": "f1", "rating": 1499.5, "game-mode": "5v5"}))  # default group (diamond)
    broker.publish(ex, q, json.dumps({"id": "f2", "detail": {"rating": 2100}, "game-mode": "5v5"}))
    broker.deliver_all()
    assert len(broker.acked) == 2 and pool.stats["enqueued"] == 2  # max_batch = 2: flushed before any tick
    groups = {pkg.synth.REFERENCE_GROUP_NAMES[oracle_group(cfg, r)] for r in eng.pool_read()["rating"].tolist()}
    assert groups == {"diamond", "gold"}
    for i in range(4):                                                                       # capacity is 4
        broker.publish(ex, q, json.dumps({"id": f"c{i}", "rating": 2100 + i, "game-mode": "5v5"}))
    broker.deliver_all()
    pool.flush()
    assert pool.stats["enqueued"] == 4 and pool.stats["invalid"] == 2 and len(broker.nacked) == 5
    assert pool.tick() == 0 and pool.in_queue("c0") and not pool.in_queue("c3")
    assert pool.remove_user("nobody") == ("ok", "removed")  # ActiveUser.remove_user/1 is idempotent
    ok, st = next(iter(pool.workers.values())).status()
    assert ok == "ok" and st["pool"]["message_count"] == 4


def oracle_group(cfg, rating):
    from oracle import oracle as orc
    return orc.find_rating_group(cfg, rating)
