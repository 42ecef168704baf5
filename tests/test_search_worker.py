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
    pool = sw.SearchPool(eng, ["1v1", "5v5"], pkg.synth.REFERENCE_GROUP_NAMES)
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
    assert workers["gold"].config["qos"] == {"prefetch_count": 10}
    players = [("u1", 100, "1v1"), ("u2", 2100, "1v1"), ("u3", 150, "1v1"), ("u4", 2200, "1v1"),
               ("u5", 4500, "1v1"), ("u1", 100, "1v1")]  # u1 twice: "already in the queue"
    for pid, r, m in players:
        publish_player(pkg, broker, cfg, pid, r, m)
    assert broker.deliver_all() == 6
    assert not broker.acked  # nothing is acked before the players are resident
    assert pool.tick() == 2
    assert len(broker.acked) == 6 and not broker.nacked
    assert pool.stats == {"enqueued": 5, "duplicates": 1, "invalid": 0, "lobbies": 2}
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
                         window=sw.WindowSchedule(w0=10, growth_per_s=20, w_max=400), clock=lambda: now[0])
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
    assert pool.in_queue("e") and not pool.enqueued_at.keys() - {sw.player_handle("e")}
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


def test_player_handle_is_stable_and_in_range():
    h = sw.player_handle("c0a8012e-1c9b-4b7e-9d2f-5f1d3a2b4c6d")
    assert h == sw.player_handle("c0a8012e-1c9b-4b7e-9d2f-5f1d3a2b4c6d") and 0 <= h < 2 ** 64 - 2
    assert len({sw.player_handle(f"p{i}") for i in range(10000)}) == 10000


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
