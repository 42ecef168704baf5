"""EXTENSION beyond the reference: strategist policy S1 — a lobby may span at most W
rating points.  CPU half: the oracle's windowed walk (orc_run_windowed) against an independent
pure-Python restatement of its definition.  GPU half: the CUDA tick with mm_set_option("max_spread", W)
against that oracle, bit-exact (lobbies, member order, leftovers in enqueue order)."""
import numpy as np
import pytest

ARRIVAL, RATING = 0, 1


def py_windowed(cfg, W, ids, rating, mode, alive=None):
    """Definition, straight from oracle/mm_oracle.h: per (mode, group) partition sorted by (clamp key, seq):
    i = 0; while i + L <= n: spread(i .. i+L-1) <= W ? emit, i += L : player i stays queued, i += 1."""
    G = cfg.n_groups
    rmin = min(cfg.group_lo[g] for g in range(G))
    rmax = max(cfg.group_hi[g] for g in range(G))
    parts = {}
    for i, (pid, r, m) in enumerate(zip(ids, rating, mode)):
        if alive is not None and not alive[i]:
            continue
        g = cfg.default_group
        for k in range(G):
            if cfg.group_lo[k] <= r <= cfg.group_hi[k]:
                g = k
                break
        key = min(max(int(r), rmin - 1), rmax + 1)
        parts.setdefault((int(m), g), []).append((key, i, int(pid)))
    lobbies, resid = [], []
    for (m, g) in sorted(parts):
        p = sorted(parts[(m, g)])
        L = cfg.modes[m].teams * cfg.modes[m].team_size
        i = 0
        while i + L <= len(p):
            if W < 0 or p[i + L - 1][0] - p[i][0] <= W:
                lobbies.append((m, g, tuple(x[2] for x in p[i:i + L])))
                i += L
            else:
                resid.append(p[i][1])
                i += 1
        resid.extend(x[1] for x in p[i:])
    return lobbies, [int(ids[i]) for i in sorted(resid)]


def oracle_lobbies(ref):
    return [(int(h["mode"]), int(h["group"]), tuple(int(x) for x in ref.member_ids[h["first_member"]:h["first_member"] + h["n_members"]]))
            for h in ref.lobbies]


def small_pool(pkg, seed, n, n_modes=2, lo=-20, hi=5020):
    rng = np.random.default_rng(seed)
    ids = pkg.synth.mix64(np.arange(n, dtype=np.uint64) + np.uint64(seed * 1000003))
    rating = rng.integers(lo, hi, n).astype(np.int32)
    mode = rng.integers(0, n_modes, n).astype(np.uint8)
    ts = np.arange(n, dtype=np.uint32)
    return ids, rating, mode, ts


# ------------------------------------------------------------------------------- CPU: the oracle
def test_windowed_known_answer(pkg, oracle):
    cfg = pkg.synth.make_config(n_groups=1, modes=(("1v1", 2, 1),), order=RATING, default_group=0)
    ids = np.arange(1, 9, dtype=np.uint64)
    rating = np.array([100, 120, 400, 430, 460, 900, 901, 902], np.int32)
    r = oracle.run_windowed(cfg, 50, ids, rating, np.zeros(8, np.uint8))
    assert [x[2] for x in oracle_lobbies(r)] == [(1, 2), (3, 4), (6, 7)]
    assert list(r.residual_ids) == [5, 8]
    r0 = oracle.run_windowed(cfg, 0, ids, rating, np.zeros(8, np.uint8))
    assert r0.n_lobbies == 0 and list(r0.residual_ids) == list(ids)


@pytest.mark.parametrize("W", [-1, 0, 1, 7, 60, 100000])
@pytest.mark.parametrize("n", [0, 1, 5, 64, 700, 5000])
def test_windowed_oracle_equals_definition(pkg, oracle, n, W):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=RATING)
    lo, hi = (-20, 5020) if n >= 700 else (1400, 1600)  # small pools: dense ratings so that windows fill
    ids, rating, mode, _ = small_pool(pkg, 3 + n, n, lo=lo, hi=hi)
    alive = (np.random.default_rng(n).random(n) > 0.1).astype(np.uint8)
    r = oracle.run_windowed(cfg, W, ids, rating, mode, alive)
    lob, resid = py_windowed(cfg, W, ids, rating, mode, alive)
    assert oracle_lobbies(r) == lob
    assert [int(x) for x in r.residual_ids] == resid
    assert r.n_dead == int((alive == 0).sum())


def test_unlimited_window_is_the_reference_policy(pkg, oracle):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=RATING)
    ids, rating, mode, _ = small_pool(pkg, 9, 20000)
    a, b = oracle.run_windowed(cfg, -1, ids, rating, mode), oracle.run_literal(cfg, ids, rating, mode)
    assert np.array_equal(a.lobbies, b.lobbies) and np.array_equal(a.member_ids, b.member_ids)
    assert np.array_equal(a.residual_ids, b.residual_ids)


def test_every_lobby_respects_the_window(pkg, oracle):
    cfg = pkg.synth.make_config(n_groups=8, order=RATING)
    ids, rating, mode, _ = small_pool(pkg, 17, 30000, lo=0, hi=5001)
    by_id = dict(zip(ids.tolist(), rating.tolist()))
    for W in (0, 3, 25):
        r = oracle.run_windowed(cfg, W, ids, rating, mode)
        for _, _, members in oracle_lobbies(r):
            rs = [by_id[m] for m in members]
            assert max(rs) - min(rs) <= W
        assert r.n_matched + r.n_residual == len(ids)


# ------------------------------------------------------------------------------- GPU: parity
def check(eng, ref, lob, mem, seq, st, seq_of=None):
    assert (st.n_lobbies, st.n_matched, st.n_residual, st.n_dead) == (ref.n_lobbies, ref.n_matched, ref.n_residual, ref.n_dead)
    assert np.array_equal(lob, ref.lobbies)
    assert np.array_equal(mem, ref.member_ids)
    if seq is not None:
        assert np.array_equal(seq, ref.emit_seq if seq_of is None else np.asarray(seq_of, np.uint32)[ref.emit_seq])
    assert np.array_equal(eng.pool_read()["id"], ref.residual_ids)


@pytest.mark.gpu
@pytest.mark.parametrize("wide", [0, 1])  # 0: wide rating groups stored as several partitions — windows cross them
@pytest.mark.parametrize("tick_impl", [1, 0])
@pytest.mark.parametrize("W", [-1, 0, 2, 40, 100000])
@pytest.mark.parametrize("n", [0, 3, 2049, 70001, 600_011])
def test_window_parity_random_pools(pkg, oracle, n, W, tick_impl, wide):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=RATING, capacity=max(n, 1),
                                flags=pkg.abi.MM_F_WIDE_PARTITIONS * wide)
    ids, rating, mode, ts = small_pool(pkg, 21 + n, n)
    alive = (np.random.default_rng(n + 1).random(n) > 0.05).astype(np.uint8)
    with pkg.Engine(cfg) as eng:
        eng.set_option("tick_impl", tick_impl)
        eng.set_option("max_spread", W)
        assert eng.enqueue(ids, rating, mode, ts).all()
        eng.remove(ids[alive == 0])
        lob, mem, seq, st = eng.tick()
        check(eng, oracle.run_windowed(cfg, W, ids, rating, mode, alive), lob, mem, seq, st)


@pytest.mark.gpu
@pytest.mark.parametrize("W", [0, 1, 4])
def test_window_parity_few_bins_both_rankings(pkg, oracle, W):
    """A 41-value rating domain: partitions of ~20 bins, so the ballot tile sort and its staged write-back run in
    RATING order with a large share of the pool staying queued; the hashed lists (heavy bins) must agree."""
    n = 300_007
    cfg = pkg.synth.make_config(groups=((0, 19), (20, 39)), modes=(("1v1", 2, 1), ("3v3", 2, 3)), order=RATING,
                                capacity=n, default_group=1)
    ids, rating, mode, ts = small_pool(pkg, 5, n, lo=-3, hi=44)
    for rank_impl in (3, 2):
        with pkg.Engine(cfg) as eng:
            eng.set_option("rank_impl", rank_impl)
            eng.set_option("max_spread", W)
            assert eng.enqueue(ids, rating, mode, ts).all()
            lob, mem, seq, st = eng.tick()
            check(eng, oracle.run_windowed(cfg, W, ids, rating, mode), lob, mem, seq, st)


@pytest.mark.gpu
def test_window_sparse_pool_leaves_most_players_queued(pkg, oracle):
    """5v5 with W=1 on a sparse pool: most seeds fail, leftovers are ~the whole pool (stresses the staged compaction)."""
    n = 400_003
    cfg = pkg.synth.make_config(n_groups=32, modes=(("5v5", 2, 5),), order=RATING, capacity=n)
    ids, rating, mode, ts = small_pool(pkg, 77, n, n_modes=1, lo=0, hi=5001)
    with pkg.Engine(cfg) as eng:
        eng.set_option("max_spread", 0)
        assert eng.enqueue(ids, rating, mode, ts).all()
        lob, mem, seq, st = eng.tick()
        ref = oracle.run_windowed(cfg, 0, ids, rating, mode)
        check(eng, ref, lob, mem, seq, st)
        assert st.n_residual > 0 and st.n_matched > 0


@pytest.mark.gpu
def test_window_multi_tick_accumulation(pkg, oracle):
    """Leftovers stay resident in enqueue order; later arrivals complete their windows (what the reference's
    requeue + save_new_state would do across messages)."""
    cfg = pkg.synth.make_config(n_groups=8, order=RATING, capacity=400_000)
    W = 3
    with pkg.Engine(cfg) as eng:
        eng.set_option("max_spread", W)
        q_ids = np.zeros(0, np.uint64); q_r = np.zeros(0, np.int32); q_m = np.zeros(0, np.uint8)
        q_seq = np.zeros(0, np.uint32)
        for t in range(5):
            ids, rating, mode, ts = small_pool(pkg, 1000 + t, 60_000, lo=0, hi=5001)
            ids = ids + np.uint64(t) * np.uint64(1 << 40)
            assert eng.enqueue(ids, rating, mode, ts).all()
            q_ids = np.concatenate([q_ids, ids]); q_r = np.concatenate([q_r, rating]); q_m = np.concatenate([q_m, mode])
            q_seq = np.concatenate([q_seq, (t * 60_000 + np.arange(60_000)).astype(np.uint32)])
            lob, mem, seq, st = eng.tick()
            ref = oracle.run_windowed(cfg, W, q_ids, q_r, q_m)
            check(eng, ref, lob, mem, seq, st, seq_of=q_seq)
            keep = np.isin(q_ids, ref.residual_ids)
            q_ids, q_r, q_m, q_seq = q_ids[keep], q_r[keep], q_m[keep], q_seq[keep]
            assert eng.in_queue(q_ids[:100]).all()
        # switching the policy off matches the rest with the reference rule
        eng.set_option("max_spread", -1)
        lob, mem, seq, st = eng.tick()
        check(eng, oracle.run_literal(cfg, q_ids, q_r, q_m), lob, mem, seq, st, seq_of=q_seq)


@pytest.mark.gpu
@pytest.mark.parametrize("n_modes", [5, 8])
@pytest.mark.parametrize("W", [-1, 30])
def test_large_key_domains_use_the_fallback_tail_layouts(pkg, oracle, n_modes, W):
    """5 x 5003 bins: the scan tail reads the bin keys from global memory; 8 x 5003 bins: it also parks the matched
    counts there (and the placement runs one CTA per SM with a shallower ring).  Both policies."""
    shapes = (("1v1", 2, 1), ("2v2", 2, 2), ("3v3", 2, 3), ("5v5", 2, 5), ("solo4", 4, 1), ("duo3", 3, 2), ("6v6", 2, 6),
              ("solo3", 3, 1))[:n_modes]
    n = 120_011
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, modes=shapes, order=RATING, capacity=n)
    ids, rating, mode, ts = small_pool(pkg, 31 + n_modes, n, n_modes=n_modes)
    with pkg.Engine(cfg) as eng:
        eng.set_option("max_spread", W)
        assert eng.enqueue(ids, rating, mode, ts).all()
        lob, mem, seq, st = eng.tick()
        check(eng, oracle.run_windowed(cfg, W, ids, rating, mode), lob, mem, seq, st)


@pytest.mark.gpu
def test_window_many_partitions(pkg, oracle):
    """64 groups x 4 modes = 256 partitions: every warp of the tail walks several pairs of partitions."""
    shapes = (("1v1", 2, 1), ("2v2", 2, 2), ("3v3", 2, 3), ("5v5", 2, 5))
    n = 400_009
    cfg = pkg.synth.make_config(n_groups=64, modes=shapes, order=RATING, capacity=n)
    ids, rating, mode, ts = small_pool(pkg, 64, n, n_modes=4, lo=0, hi=5001)
    for W in (0, 6):
        with pkg.Engine(cfg) as eng:
            eng.set_option("max_spread", W)
            assert eng.enqueue(ids, rating, mode, ts).all()
            lob, mem, seq, st = eng.tick()
            check(eng, oracle.run_windowed(cfg, W, ids, rating, mode), lob, mem, seq, st)


@pytest.mark.gpu
def test_window_rejected_in_arrival_order(pkg):
    cfg = pkg.synth.make_config(n_groups=4, order=ARRIVAL, capacity=1000)
    with pkg.Engine(cfg) as eng:
        with pytest.raises(Exception):
            eng.set_option("max_spread", 10)
        eng.set_option("max_spread", -1)


@pytest.mark.gpu
def test_window_ten_million(pkg, oracle):
    w = pkg.synth.WORKLOADS["config3_10m_g32_5v5"]
    cfg, m = pkg.synth.workload_config("config3_10m_g32_5v5", RATING, w["n"])
    ids, rating, mode, ts = pkg.synth.gen_pool(3, w["n"], mode=m)
    with pkg.Engine(cfg) as eng:
        eng.set_option("max_spread", 0)
        assert eng.enqueue(ids, rating, mode, ts).all()
        lob, mem, seq, st = eng.tick()
        ref = oracle.run_windowed(cfg, 0, ids, rating, mode)
        check(eng, ref, lob, mem, seq, st)
