"""Streaming ingest (BASELINE.json configs[4], one GPU): Poisson arrivals at `rate` players/s, one search
tick every `dt_ms`, p50/p99/p99.9 of the per-player queue->match latency

    latency = t(host holds the lobby that contains the player) - t(player arrived)

measured in real time (the loop is paced with the wall clock; a tick that overruns its period is counted).
Strict-parity mode: there are no time-expanded windows, so a player waits only until L-1 more
players of its (mode, group) partition have arrived and the next tick fires.
With max_spread=W (extension, policy S1) a lobby may span at most W rating points, so the pool accumulates the
players whose neighbourhood is still sparse and the latency distribution grows a tail.
usage: python tools/stream_bench.py [rate=1e6] [seconds=2] [dt_ms=1,5] [groups=32] [mode=5v5] [max_spread=-1]"""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

pkg = importlib.import_module("microservice-matchmaking_b200")
args = dict(a.split("=") for a in sys.argv[1:])
rate = float(args.get("rate", 1e6)); seconds = float(args.get("seconds", 2.0))
dts = [float(x) for x in args.get("dt_ms", "1,5").split(",")]
G = int(args.get("groups", 32)); mode_name = args.get("mode", "5v5")
modes = (("1v1", 2, 1),) if mode_name == "1v1" else (("5v5", 2, 5),)
L = modes[0][1] * modes[0][2]
W = int(args.get("max_spread", -1))

for dt_ms in dts:
    dt = dt_ms * 1e-3
    n_total = int(rate * seconds)
    rng = np.random.default_rng(1)
    arrive = np.cumsum(rng.exponential(1.0 / rate, n_total))          # Poisson process
    ids, rating, _, _ = pkg.synth.gen_pool(3, n_total)
    mode = np.zeros(n_total, np.uint8)
    cfg = pkg.synth.make_config(n_groups=G, modes=modes, order=pkg.abi.MM_ORDER_RATING, capacity=1 << 20,
                                active_capacity=4 * n_total)
    eng = pkg.Engine(cfg)
    eng.set_option("max_spread", W)
    eng.enqueue(ids[:10], rating[:10], mode[:10]); eng.tick(); eng.remove(ids[:10])   # warm-up / allocation
    matched_at = np.full(n_total, np.nan)
    order = np.argsort(ids, kind="stable"); sorted_ids = ids[order]
    lo, k, overruns, tick_us, per_tick = 10, 0, 0, [], []
    t0 = time.perf_counter()
    n_ticks = int(seconds / dt)
    for k in range(1, n_ticks + 1):
        deadline = t0 + k * dt
        while time.perf_counter() < deadline:
            pass
        now = time.perf_counter() - t0
        hi = int(np.searchsorted(arrive, now))                         # everyone who has arrived by now
        if hi > lo:
            eng.enqueue(ids[lo:hi], rating[lo:hi], mode[lo:hi], (arrive[lo:hi] * 1e6).astype(np.uint32))
        lob, mem, _, st = eng.tick(want_emit_seq=False)
        done = time.perf_counter() - t0
        if len(mem):
            idx = order[np.searchsorted(sorted_ids, mem)]
            matched_at[idx] = done
        tick_us.append(st.device_us); per_tick.append(hi - lo)
        if done > (k + 1) * dt:
            overruns += 1
        lo = max(lo, hi)
    lat = (matched_at - arrive)[10:lo]
    ok = ~np.isnan(lat)
    q = lambda p: float(np.percentile(lat[ok], p) * 1e3)
    print(json.dumps({"workload": "stream", "rate_per_s": rate, "dt_ms": dt_ms, "seconds": seconds, "groups": G, "mode": mode_name, "max_spread": W,
                      "players_enqueued": int(lo - 10), "matched": int(ok.sum()), "still_queued": int((~ok).sum()),
                      "latency_ms": {"p50": q(50), "p99": q(99), "p99.9": q(99.9), "max": q(100)},
                      "players_per_tick_mean": float(np.mean(per_tick)), "tick_device_us_mean": float(np.mean(tick_us)),
                      "tick_device_us_p99": float(np.percentile(tick_us, 99)), "ticks": n_ticks, "overrun_ticks": overruns}), flush=True)
    eng.close()
