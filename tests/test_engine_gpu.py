"""GPU parity tests: the CUDA search tick, called through the C ABI, against the CPU
oracle on the same seeded inputs — bit-exact (integer / index work, no tolerance)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ARRIVAL, RATING = 0, 1


def lobby_lists(lob, mem):
    return [tuple(mem[h["first_member"]:h["first_member"] + h["n_members"]]) for h in lob]


def assert_tick_matches(eng, ref, lob, mem, seq, st, seq_of=None):
    """seq_of[i] = enqueue sequence number of the oracle's i-th input player (identity for a fresh engine whose
    every offered player was fed to the oracle): the engine reports emission order in sequence numbers."""
    assert st.n_lobbies == ref.n_lobbies
    assert st.n_matched == ref.n_matched
    assert st.n_residual == ref.n_residual
    assert st.n_dead == ref.n_dead
    assert np.array_equal(lob, ref.lobbies)
    assert np.array_equal(mem, ref.member_ids)
    if seq is not None:
        assert np.array_equal(seq, ref.emit_seq if seq_of is None else np.asarray(seq_of, np.uint32)[ref.emit_seq])
    assert np.array_equal(eng.pool_read()["id"], ref.residual_ids)


def make_pool(pkg, seed, n, n_modes=2, bell=False, oor=0.01):
    rng = np.random.default_rng(seed)
    ids, rating, _, ts = pkg.synth.gen_pool(seed, n, bell=bell)
    rating = rating.copy()
    k = int(n * oor)
    if k:
        rating[rng.integers(0, n, k)] = rng.integers(-100, 5200, k)
    mode = rng.integers(0, n_modes, n).astype(np.uint8)
    return ids, rating, mode, ts


@pytest.mark.parametrize("wide", [0, 1])  # 1: MM_F_WIDE_PARTITIONS — a 1 500-rating group stays ONE partition (list ranking)
@pytest.mark.parametrize("impl", [3, 2])
@pytest.mark.parametrize("order", [ARRIVAL, RATING])
@pytest.mark.parametrize("n", [0, 1, 2, 9, 31, 33, 1000, 2047, 2048, 2049, 4097, 70001])
def test_random_pool_matches_literal_oracle(pkg, oracle, n, order, impl, wide):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=order, capacity=max(n, 1),
                                flags=pkg.abi.MM_F_WIDE_PARTITIONS * wide)
    ids, rating, mode, ts = make_pool(pkg, 11 + n, n)
    with pkg.Engine(cfg) as eng:
        eng.set_option("rank_impl", impl)
        acc = eng.enqueue(ids, rating, mode, ts)
        assert (acc == 1).all()
        assert eng.pool_size() == n
        lob, mem, seq, st = eng.tick()
        ref = oracle.run_literal(cfg, ids, rating, mode)
        assert_tick_matches(eng, ref, lob, mem, seq, st)
        if order == ARRIVAL and n:
            # sorting by emit_seq reproduces the serialized reference's emission order
            assert np.array_equal(np.argsort(seq, kind="stable"), np.argsort(ref.emission_rank, kind="stable"))


@pytest.mark.parametrize("wide", [0, 1])
@pytest.mark.parametrize("order", [ARRIVAL, RATING])
@pytest.mark.parametrize("bell", [False, True])
def test_one_million_mixed(pkg, oracle, order, bell, wide):
    n = 1_000_003
    cfg = pkg.synth.make_config(n_groups=8, order=order, capacity=n, flags=pkg.abi.MM_F_WIDE_PARTITIONS * wide)
    ids, rating, mode, ts = make_pool(pkg, 5, n, bell=bell, oor=0.0)
    with pkg.Engine(cfg) as eng:
        assert eng.enqueue(ids, rating, mode, ts).all()
        lob, mem, seq, st = eng.tick()
        ref = oracle.run_closed_form(cfg, ids, rating, mode)
        assert_tick_matches(eng, ref, lob, mem, seq, st)


def test_config2_and_both_rank_impls_agree(pkg, oracle):
    # BASELINE.json configs[1]: 1M players, 8 rating groups, 1v1
    w = pkg.synth.WORKLOADS["config2_1m_g8_1v1"]
    cfg = pkg.synth.make_config(n_groups=w["n_groups"], order=RATING, capacity=w["n"])
    ids, rating, mode, ts = pkg.synth.gen_pool(1, w["n"], mode=w["mode"])
    out = []
    for impl in (3, 2):
        with pkg.Engine(cfg) as eng:
            eng.set_option("rank_impl", impl)
            assert eng.enqueue(ids, rating, mode, ts).all()
            lob, mem, seq, st = eng.tick()
            out.append((lob, mem, seq))
    assert all(np.array_equal(a, b) for a, b in zip(out[0], out[1]))
    ref = oracle.run_closed_form(cfg, ids, rating, mode)
    assert np.array_equal(out[0][1], ref.member_ids) and np.array_equal(out[0][0], ref.lobbies)


@pytest.mark.parametrize("order", [ARRIVAL, RATING])
@pytest.mark.parametrize("n", [0, 1, 7, 2047, 2049, 4097, 300_001, 2_000_003])
def test_fused_single_launch_equals_split_and_oracle(pkg, oracle, n, order):
    """k_tick (one cooperative launch) vs the four-kernel tick vs the oracle."""
    cfg = pkg.synth.make_config(n_groups=8, order=order, capacity=max(n, 1) + 100)
    ids, rating, mode, ts = make_pool(pkg, 100 + n, n, bell=(n % 2 == 1))
    rng = np.random.default_rng(n)
    alive = (rng.random(n) > 0.03).astype(np.uint8)
    outs = []
    for impl in (1, 0):
        with pkg.Engine(cfg) as eng:
            eng.set_option("tick_impl", impl)
            assert eng.enqueue(ids, rating, mode, ts).all()
            eng.remove(ids[alive == 0])
            lob, mem, seq, st = eng.tick()
            if impl == 1:
                assert st.n_launches == 1, "the fused cooperative kernel was not used"
            ref = oracle.run_closed_form(cfg, ids, rating, mode, alive)
            assert_tick_matches(eng, ref, lob, mem, seq, st)
            # second tick on the compacted pool + new arrivals
            ids2, rating2, _, ts2 = pkg.synth.gen_pool(999, 5000, first=10 ** 9)
            mode2 = (np.arange(5000) % 2).astype(np.uint8)
            room = cfg.capacity - eng.pool_size()
            k = min(5000, room)
            assert eng.enqueue(ids2[:k], rating2[:k], mode2[:k], ts2[:k]).all()
            st2 = eng.tick_device()
            outs.append((lob, mem, seq, st2.n_lobbies, st2.n_matched, eng.pool_read()["id"]))
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))


@pytest.mark.parametrize("rank_impl", [3, 2])
@pytest.mark.parametrize("n_groups,n_modes", [(7, 2), (32, 1), (32, 2), (45, 2), (64, 4)])
def test_arrival_small_key_domain_variants(pkg, oracle, rank_impl, n_groups, n_modes):
    """ARRIVAL order has one bin per (mode, group) partition: the ballot tile sort (3) degenerates to a stable
    compaction of the removed players; the hashed lists (2) must agree."""
    modes = (("1v1", 2, 1), ("5v5", 2, 5), ("2v2", 2, 2), ("3v3", 2, 3))[:n_modes]
    n = 150_001
    cfg = pkg.synth.make_config(n_groups=n_groups, modes=modes, order=ARRIVAL, capacity=n)
    ids, rating, mode, ts = make_pool(pkg, 7 * n_groups + n_modes, n, n_modes=n_modes, oor=0.0, bell=True)
    rng = np.random.default_rng(1)
    alive = (rng.random(n) > 0.02).astype(np.uint8)
    for tick_impl in (1, 0):
        with pkg.Engine(cfg) as eng:
            eng.set_option("rank_impl", rank_impl)
            eng.set_option("tick_impl", tick_impl)
            assert eng.enqueue(ids, rating, mode, ts).all()
            eng.remove(ids[alive == 0])
            lob, mem, seq, st = eng.tick()
            assert_tick_matches(eng, oracle.run_closed_form(cfg, ids, rating, mode, alive), lob, mem, seq, st)


def test_degenerate_everyone_same_rating(pkg, oracle):
    n = 200_000
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=RATING, capacity=n)
    ids, _, _, ts = pkg.synth.gen_pool(3, n)
    rating = np.full(n, 1500, np.int32)
    mode = np.ones(n, np.uint8)
    with pkg.Engine(cfg) as eng:
        assert eng.enqueue(ids, rating, mode, ts).all()
        lob, mem, seq, st = eng.tick()
        ref = oracle.run_closed_form(cfg, ids, rating, mode)
        assert_tick_matches(eng, ref, lob, mem, seq, st)
        assert np.array_equal(mem, ids[: (n // 10) * 10])  # pure enqueue order


def test_few_distinct_ratings(pkg, oracle):
    # ratings rounded to 10: every bin is shared by several warps in every round
    n = 300_000
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=RATING, capacity=n)
    ids, rating, mode, ts = make_pool(pkg, 9, n, oor=0.0)
    rating = (rating // 10 * 10).astype(np.int32)
    with pkg.Engine(cfg) as eng:
        assert eng.enqueue(ids, rating, mode, ts).all()
        lob, mem, seq, st = eng.tick()
        assert_tick_matches(eng, oracle.run_closed_form(cfg, ids, rating, mode), lob, mem, seq, st)


def test_overlapping_unsorted_groups(pkg, oracle):
    groups = [(3000, 5000), (0, 1200), (1000, 3500)]  # first match in list order; default = none
    n = 50_000
    for order in (ARRIVAL, RATING):
        cfg = pkg.synth.make_config(groups=groups, order=order, capacity=n, default_group=1)
        ids, rating, mode, ts = make_pool(pkg, 21, n, oor=0.02)
        with pkg.Engine(cfg) as eng:
            assert eng.enqueue(ids, rating, mode, ts).all()
            lob, mem, seq, st = eng.tick()
            assert_tick_matches(eng, oracle.run_literal(cfg, ids, rating, mode), lob, mem, seq, st)


def test_many_modes_and_team_shapes(pkg, oracle):
    modes = (("1v1", 2, 1), ("2v2", 2, 2), ("3v3v3", 3, 3), ("5v5", 2, 5), ("solo7", 7, 1))
    n = 120_000
    for order in (ARRIVAL, RATING):
        cfg = pkg.synth.make_config(n_groups=16, modes=modes, order=order, capacity=n)
        ids, rating, mode, ts = make_pool(pkg, 33, n, n_modes=len(modes), oor=0.0)
        with pkg.Engine(cfg) as eng:
            assert eng.enqueue(ids, rating, mode, ts).all()
            lob, mem, seq, st = eng.tick()
            assert_tick_matches(eng, oracle.run_literal(cfg, ids, rating, mode), lob, mem, seq, st)
            assert set(np.unique(lob["n_members"])) <= {2, 4, 9, 10, 7}


# ---- active set: dedupe / remove / in_queue (models/active_user.ex, middleware/worker.ex:65-70) ----
def test_dedupe_within_and_across_batches(pkg):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=ARRIVAL, capacity=100)
    with pkg.Engine(cfg) as eng:
        acc = eng.enqueue([5, 6, 5, 7, 6, 5], [100] * 6, [0] * 6)
        assert list(acc) == [1, 1, 0, 1, 0, 0]  # first occurrence wins
        assert list(eng.pool_read()["id"]) == [5, 6, 7]
        acc = eng.enqueue([8, 5, 9], [100, 100, 100], [0, 0, 0])
        assert list(acc) == [1, 0, 1]  # "You are already in the queue."
        assert list(eng.pool_read()["id"]) == [5, 6, 7, 8, 9]
        assert list(eng.in_queue([5, 9, 10])) == [True, True, False]
        assert eng.active_size() == 5


def test_invalid_players_rejected(pkg):
    cfg = pkg.synth.make_config(n_groups=2, order=ARRIVAL, capacity=100)  # G=2: no default group
    assert cfg.default_group == -1
    with pkg.Engine(cfg) as eng:
        acc = eng.enqueue([1, 2, 3, 4, 2 ** 64 - 1], [100, 6000, -5, 100, 100], [0, 0, 0, 7, 0])
        assert list(acc) == [1, 2, 2, 2, 2]
        assert list(eng.pool_read()["id"]) == [1]


def test_pool_capacity(pkg):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=ARRIVAL, capacity=5)
    with pkg.Engine(cfg) as eng:
        acc = eng.enqueue(np.arange(1, 9), [100] * 8, [1] * 8)
        assert list(acc) == [1, 1, 1, 1, 1, 3, 3, 3]
        assert eng.pool_size() == 5 and eng.active_size() == 5
        assert not eng.in_queue([6])[0]


def test_matched_players_stay_active_until_removed(pkg):
    # the lobby stage removes them later (game-lobby/worker.ex:80); search only emits
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=ARRIVAL, capacity=16)
    with pkg.Engine(cfg) as eng:
        eng.enqueue([1, 2, 3], [100, 100, 100], [0, 0, 0])
        lob, mem, _, st = eng.tick()
        assert lobby_lists(lob, mem) == [(1, 2)] and eng.pool_size() == 1
        assert list(eng.in_queue([1, 2, 3])) == [True, True, True]
        assert list(eng.enqueue([1], [100], [0])) == [0]  # still "in the queue"
        assert eng.remove([1, 2, 99]) == 2
        assert list(eng.in_queue([1, 2, 3])) == [False, False, True]
        assert list(eng.enqueue([1], [100], [0])) == [1]
        lob, mem, _, st = eng.tick()
        assert lobby_lists(lob, mem) == [(3, 1)]


@pytest.mark.parametrize("order", [ARRIVAL, RATING])
def test_leavers_are_filtered(pkg, oracle, order):
    # search/worker.ex:267-280: players removed while queued never reach a lobby
    n = 60_000
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=order, capacity=n)
    ids, rating, mode, ts = make_pool(pkg, 41, n)
    rng = np.random.default_rng(4)
    alive = (rng.random(n) > 0.1).astype(np.uint8)
    with pkg.Engine(cfg) as eng:
        assert eng.enqueue(ids, rating, mode, ts).all()
        gone = ids[alive == 0]
        assert eng.remove(np.concatenate([gone, gone[:10]])) == len(gone)  # twins in a batch count once
        assert not eng.in_queue(gone).any()
        lob, mem, seq, st = eng.tick()
        ref = oracle.run_literal(cfg, ids, rating, mode, alive)
        assert_tick_matches(eng, ref, lob, mem, seq, st)
        assert not np.isin(gone, mem).any()


def test_kat_leaver(pkg):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=ARRIVAL, capacity=16)
    with pkg.Engine(cfg) as eng:
        eng.enqueue([1, 2, 3, 4, 5], [100] * 5, [0] * 5)
        eng.remove([2])
        lob, mem, seq, st = eng.tick()
        assert lobby_lists(lob, mem) == [(1, 3), (4, 5)] and st.n_dead == 1 and list(seq) == [2, 4]


# ---- residual players stay queued, in order, across ticks (save_new_state analogue) ----
@pytest.mark.parametrize("order", [ARRIVAL, RATING])
def test_multi_tick_stream(pkg, oracle, order):
    cfg = pkg.synth.make_config(n_groups=8, order=order, capacity=40_000)
    rng = np.random.default_rng(8)
    queued = [np.zeros(0, np.uint64), np.zeros(0, np.int32), np.zeros(0, np.uint8)]
    alive = np.zeros(0, np.uint8)
    qseq = np.zeros(0, np.uint32)  # enqueue sequence number of every queued player
    with pkg.Engine(cfg) as eng:
        first = 0
        for step in range(6):
            n = int(rng.integers(1, 9000))
            ids, rating, _, ts = pkg.synth.gen_pool(77, n, first=first)
            first += n
            mode = rng.integers(0, 2, n).astype(np.uint8)
            assert eng.enqueue(ids, rating, mode, ts).all()
            queued = [np.concatenate([q, x]) for q, x in zip(queued, (ids, rating, mode))]
            qseq = np.concatenate([qseq, (first - n + np.arange(n)).astype(np.uint32)])
            alive = np.concatenate([alive, np.ones(n, np.uint8)])
            lob, mem, seq, st = eng.tick()
            ref = oracle.run_literal(cfg, *queued, alive=alive)
            assert_tick_matches(eng, ref, lob, mem, seq, st, seq_of=qseq)
            keep = np.isin(queued[0], ref.residual_ids)
            queued = [q[keep] for q in queued]
            qseq = qseq[keep]
            alive = np.ones(len(queued[0]), np.uint8)
            pr = eng.pool_read()
            assert np.array_equal(pr["id"], queued[0]) and np.array_equal(pr["rating"], queued[1])
            assert np.array_equal(pr["mode"], queued[2])
            assert np.array_equal(pr["team_size"], np.where(queued[2] == 0, 1, 5))
            # a residual player can still be removed after the pool was compacted
            if len(queued[0]):
                assert eng.remove(queued[0][:1]) == 1
                alive[0] = 0


def test_lobby_cap_too_small_consumes_nothing(pkg):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=ARRIVAL, capacity=64)
    with pkg.Engine(cfg) as eng:
        eng.enqueue(np.arange(1, 21), [100] * 20, [0] * 20)
        import ctypes as C
        lob = np.empty(3, pkg.engine.LOBBY_DTYPE)
        mem = np.empty(64, np.uint64)
        st = pkg.abi.TickStats()
        rc = eng.lib.mm_tick(eng.h, 0, lob.ctypes.data_as(C.c_void_p), 3, mem.ctypes.data_as(C.c_void_p), 64, None,
                             C.byref(st))
        assert rc == pkg.abi.MM_E_CAP and eng.pool_size() == 20
        lob, mem, _, st = eng.tick()
        assert st.n_lobbies == 10 and eng.pool_size() == 0


def test_snapshot_restore_replays_identically(pkg):
    n = 100_000
    cfg = pkg.synth.make_config(n_groups=32, order=RATING, capacity=n)
    ids, rating, mode, ts = make_pool(pkg, 17, n)
    with pkg.Engine(cfg) as eng:
        assert eng.enqueue(ids, rating, mode, ts).all()
        eng.snapshot()
        a = eng.tick()
        eng.restore()
        assert eng.pool_size() == n
        b = eng.tick()
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        # device-resident variant gives the same member list
        eng.restore()
        st = eng.tick_device()
        assert st.n_lobbies == a[3].n_lobbies and st.n_matched == a[3].n_matched


# ---- BASELINE.json full size (configs[2]): size-independent properties + numpy closed form ----
def test_config3_ten_million_5v5(pkg, oracle):
    w = pkg.synth.WORKLOADS["config3_10m_g32_5v5"]
    n, G = w["n"], w["n_groups"]
    for order in (RATING, ARRIVAL):
        cfg = pkg.synth.make_config(n_groups=G, order=order, capacity=n)
        ids, rating, mode, ts = pkg.synth.gen_pool(1, n, mode=w["mode"])
        with pkg.Engine(cfg) as eng:
            assert eng.enqueue(ids, rating, mode, ts).all()
            lob, mem, seq, st = eng.tick()
            resid = eng.pool_read()["id"]
        L = 10
        assert st.n_matched == st.n_lobbies * L and st.n_matched + st.n_residual == n
        assert st.n_residual <= (L - 1) * G
        # permutation: every queued player is in exactly one lobby or still queued
        allout = np.concatenate([mem, resid])
        assert len(allout) == n and np.array_equal(np.sort(allout), np.sort(ids))
        # every lobby is inside one rating group; rating order: sorted by (rating, enqueue)
        pos = np.argsort(ids, kind="stable")
        idx = pos[np.searchsorted(ids[pos], mem)]  # pool index of each member
        r = rating[idx].reshape(-1, L)
        los, his = pkg.synth.equal_width_groups(G)
        g = np.searchsorted(np.array(his), r[:, 0])
        assert np.array_equal(g, lob["group"])
        assert (r >= np.array(los)[g][:, None]).all() and (r <= np.array(his)[g][:, None]).all()
        # inside a group: non-decreasing rating (RATING) and, on ties, increasing enqueue index
        full = (g.repeat(L).astype(np.int64) << 50) + \
               ((rating[idx].astype(np.int64) << 25) if order == RATING else 0)
        assert (np.diff(full) >= 0).all()
        tie = np.diff(full) == 0
        assert (np.diff(idx)[tie] > 0).all()
        # exact: numpy closed form of the oracle
        lm, lg, members, res = oracle.closed_form_numpy(cfg, ids, rating, mode)
        assert np.array_equal(members, mem) and np.array_equal(res, resid)
        assert np.array_equal(lg, lob["group"]) and np.array_equal(lm, lob["mode"])


# ---- robustness of the boundary ------------------------------------------------------------------------
def test_tombstone_rehash_and_reuse(pkg):
    """Many enqueue/remove cycles on a small active set: tombstones accumulate, the table is rehashed,
    membership stays exact (models/active_user.ex semantics)."""
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=ARRIVAL, capacity=4096, active_capacity=6000)
    rng = np.random.default_rng(5)
    with pkg.Engine(cfg) as eng:
        live = set()
        for step in range(40):
            ids = rng.integers(1, 50_000, 1500).astype(np.uint64)
            acc = eng.enqueue(ids, rng.integers(0, 5001, 1500), np.zeros(1500, np.uint8))
            seen = set()
            for p, a in zip(ids.tolist(), acc.tolist()):
                if p in live or p in seen:
                    assert a == 0
                else:
                    assert a in (1, 3)
                    if a == 1:
                        seen.add(p)
            live |= seen
            lob, mem, _, st = eng.tick()
            assert set(mem.tolist()) <= live
            gone = np.array(sorted(live), np.uint64)[::2]
            assert eng.remove(gone) == len(gone)
            live -= set(gone.tolist())
            probe = rng.integers(1, 50_000, 500).astype(np.uint64)
            assert list(eng.in_queue(probe)) == [p in live for p in probe.tolist()]
            assert eng.active_size() == len(live)


def test_active_set_full_is_reported(pkg):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=ARRIVAL, capacity=1 << 16, active_capacity=1000)
    with pkg.Engine(cfg) as eng:
        with pytest.raises(pkg.EngineError) as ei:
            eng.enqueue(np.arange(1, 60_001, dtype=np.uint64), np.full(60_000, 100), np.zeros(60_000, np.uint8))
        assert ei.value.status == pkg.abi.MM_E_CAP and eng.pool_size() == 0


def test_bad_options_and_state(pkg):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=ARRIVAL, capacity=64)
    with pkg.Engine(cfg) as eng:
        for name, v in (("nope", 1), ("rank_impl", 7), ("rank_impl", 0), ("place_debug", 1)):
            with pytest.raises(pkg.EngineError):
                eng.set_option(name, v)
        with pytest.raises(pkg.EngineError) as ei:
            eng.restore()  # no snapshot yet
        assert ei.value.status == pkg.abi.MM_E_STATE


def test_external_stream(pkg, oracle):
    import torch
    n = 30_000
    cfg = pkg.synth.make_config(n_groups=8, order=RATING, capacity=n)
    ids, rating, mode, ts = make_pool(pkg, 3, n)
    s = torch.cuda.Stream()
    with pkg.Engine(cfg) as eng:
        eng.set_stream(s.cuda_stream)
        assert eng.enqueue(ids, rating, mode, ts).all()
        lob, mem, seq, st = eng.tick()
        assert_tick_matches(eng, oracle.run_closed_form(cfg, ids, rating, mode), lob, mem, seq, st)


def test_two_engines_and_threads(pkg, oracle):
    """One writer per engine, several engines per process (one per rating-group shard)."""
    import threading
    n = 40_000
    cfgs = [pkg.synth.make_config(n_groups=8, order=o, capacity=n) for o in (ARRIVAL, RATING)]
    pools = [make_pool(pkg, 50 + i, n) for i in range(2)]
    out = [None, None]

    def work(i):
        with pkg.Engine(cfgs[i]) as eng:
            assert eng.enqueue(*pools[i]).all()
            out[i] = eng.tick()

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(2):
        ref = oracle.run_closed_form(cfgs[i], *pools[i][:3])
        assert np.array_equal(out[i][1], ref.member_ids) and np.array_equal(out[i][0], ref.lobbies)


def test_large_lobbies_and_many_groups(pkg, oracle):
    modes = (("battle-royale", 1, 100), ("4x8", 4, 8))
    n = 100_000
    cfg = pkg.synth.make_config(n_groups=64, modes=modes, order=RATING, capacity=n)
    ids, rating, mode, ts = make_pool(pkg, 77, n, n_modes=2, oor=0.0)
    with pkg.Engine(cfg) as eng:
        assert eng.enqueue(ids, rating, mode, ts).all()
        lob, mem, seq, st = eng.tick()
        assert_tick_matches(eng, oracle.run_literal(cfg, ids, rating, mode), lob, mem, seq, st)
        assert set(np.unique(lob["n_members"])) <= {100, 32}


def test_async_results_complete_under_the_next_ingest(pkg, oracle):
    """mm_set_option("async_results", 1): mm_tick returns with its host copies queued; the next batch is ingested
    meanwhile; the buffers are valid after mm_results_wait and equal the blocking tick's."""
    import torch
    n = 400_000
    cfg = pkg.synth.make_config(n_groups=8, order=RATING, capacity=2 * n, active_capacity=4 * n)
    a = make_pool(pkg, 71, n)
    b_ids, b_rating, _, b_ts = pkg.synth.gen_pool(72, n, first=10 ** 9)
    b_mode = (np.arange(n) % 2).astype(np.uint8)
    lob_h = torch.empty(n, dtype=torch.int64).pin_memory()
    mem_h = torch.empty(n, dtype=torch.int64).pin_memory()
    seq_h = torch.empty(n, dtype=torch.int32).pin_memory()
    with pkg.Engine(cfg) as eng:
        eng.set_option("async_results", 1)
        assert eng.enqueue(*a).all()
        st = eng.tick_raw(lob_h.data_ptr(), n, mem_h.data_ptr(), n, seq_h.data_ptr())
        assert eng.enqueue(b_ids, b_rating, b_mode, b_ts).all()  # overlaps the copies of the tick above
        eng.results_wait()
        ref = oracle.run_closed_form(cfg, a[0], a[1], a[2])
        assert (st.n_lobbies, st.n_matched) == (ref.n_lobbies, ref.n_matched)
        lob = lob_h.numpy()[:st.n_lobbies].view(ref.lobbies.dtype)
        assert np.array_equal(lob, ref.lobbies)
        assert np.array_equal(mem_h.numpy()[:st.n_matched].view(np.uint64), ref.member_ids)
        assert np.array_equal(seq_h.numpy()[:st.n_lobbies].view(np.uint32), ref.emit_seq)
        # the second tick sees the leftovers of the first + batch B, and waits for nothing that is not there
        keep = np.isin(a[0], ref.residual_ids)
        q = [np.concatenate([x[keep], y]) for x, y in zip(a[:3], (b_ids, b_rating, b_mode))]
        qseq = np.concatenate([np.arange(n)[keep], n + np.arange(n)])
        lob2, mem2, seq2, st2 = eng.tick()
        assert_tick_matches(eng, oracle.run_closed_form(cfg, *q), lob2, mem2, seq2, st2, seq_of=qseq)
        eng.results_wait()  # no-op
        eng.set_option("async_results", 0)


# ---- packed host formats on a dense-handle engine (MM_F_DENSE_IDS; SURVEY §7.3 "dense slot index") ----------------
@pytest.mark.parametrize("order", [ARRIVAL, RATING])
def test_packed_dense_engine_equals_u64_engine_and_oracle(pkg, oracle, order):
    """mm_enqueue_packed (u32 handle + u16 mode << 13 | rating) / mm_tick_packed (u32 handles) on a direct-mapped
    active set give the lobbies of the u64 entry points and of the oracle, handle for handle."""
    n = 300_017
    cfg = pkg.synth.make_config(n_groups=32, order=order, capacity=n)
    _, rating, mode, ts = make_pool(pkg, 123, n, oor=0.0)
    handles = np.random.default_rng(9).permutation(2 * n)[:n].astype(np.uint32)  # sparse, unordered handle use
    cfgd = pkg.synth.make_config(n_groups=32, order=order, capacity=n, active_capacity=2 * n)
    cfgd.flags |= pkg.abi.MM_F_DENSE_IDS
    rng = np.random.default_rng(2)
    alive = (rng.random(n) > 0.05).astype(np.uint8)
    ref = oracle.run_closed_form(cfg, handles.astype(np.uint64), rating, mode, alive)
    with pkg.Engine(cfgd) as eng:
        acc = eng.enqueue_packed(handles, pkg.Engine.pack_key(rating, mode), ts)
        assert (acc == 1).all()
        idx, code = eng.enqueue_rejects()
        assert len(idx) == 0
        assert eng.remove_packed(handles[alive == 0]) == int((alive == 0).sum())
        assert list(eng.in_queue(handles[:64].astype(np.uint64))) == [bool(a) for a in alive[:64]]
        lob, mem, seq, st = eng.tick_packed()
        assert mem.dtype == np.uint32
        assert np.array_equal(lob, ref.lobbies) and np.array_equal(mem.astype(np.uint64), ref.member_ids)
        assert np.array_equal(seq, ref.emit_seq)
        assert np.array_equal(eng.pool_read()["id"], ref.residual_ids)
        # matched players stay active until removed; a second offer is "already in the queue"
        again = eng.enqueue_packed(handles[:1000], pkg.Engine.pack_key(rating[:1000], mode[:1000]))
        assert (again[alive[:1000] == 1] == 0).all() and (again[alive[:1000] == 0] == 1).all()


def test_packed_rejects_list(pkg):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, order=ARRIVAL, capacity=6, active_capacity=100)
    cfg.flags |= pkg.abi.MM_F_DENSE_IDS
    with pkg.Engine(cfg) as eng:
        handles = np.array([5, 6, 5, 100, 7, 8, 9, 10, 11, 12], np.uint32)  # 5 twice, 100 = out of the handle range
        mode = np.array([0, 0, 0, 0, 7, 0, 0, 0, 0, 0], np.uint8)           # mode 7 is not configured
        rating = np.full(10, 1000)
        acc = eng.enqueue_packed(handles, pkg.Engine.pack_key(rating, mode))
        assert list(acc) == [1, 1, 0, 2, 2, 1, 1, 1, 1, 3]                     # capacity 6: the last one does not fit
        idx, code = eng.enqueue_rejects()
        order = np.argsort(idx)
        assert list(idx[order]) == [2, 3, 4, 9] and list(code[order]) == [0, 2, 2, 3]
        assert eng.pool_size() == 6 and eng.active_size() == 6
        assert list(eng.in_queue(np.array([5, 12, 100], np.uint64))) == [True, False, False]
        assert eng.remove_packed(np.array([5, 5, 99], np.uint32)) == 1


def test_staged_packed_ingest_equals_the_blocking_one(pkg, oracle):
    """mm_enqueue_packed_begin / _end with two batches in flight and ticks in between == the same batches through the
    blocking mm_enqueue_packed, tick by tick (codes, reject lists, lobbies, members, emission order, leftovers); the
    blocking path itself is held to the oracle by the tests above."""
    n, steps = 120_011, 4
    cfg = pkg.synth.make_config(n_groups=32, order=RATING, capacity=2 * n, active_capacity=(steps + 1) * n)
    cfg.flags |= pkg.abi.MM_F_DENSE_IDS
    batches = []
    for k in range(steps):
        _, rating, mode, ts = make_pool(pkg, 500 + k, n, oor=0.0)
        handles = (np.arange(n, dtype=np.uint32) + np.uint32(k * n))
        if k:  # a few players of the previous batch offered again: "already in the queue"
            handles[:50] = batches[-1][0][100:150]
        batches.append((handles, rating, mode, ts))
    with pkg.Engine(cfg) as a, pkg.Engine(cfg) as b:
        with pytest.raises(pkg.EngineError):
            a.enqueue_packed_end()  # nothing staged
        a.enqueue_packed_begin(batches[0][0], pkg.Engine.pack_key(batches[0][1], batches[0][2]), batches[0][3])
        for k in range(steps):
            h, rating, mode, ts = batches[k]
            if k + 1 < steps:
                hn, rn, mn, tn = batches[k + 1]
                a.enqueue_packed_begin(hn, pkg.Engine.pack_key(rn, mn), tn)  # next upload in flight
                if k == 0:
                    with pytest.raises(pkg.EngineError):
                        a.enqueue_packed_begin(hn, pkg.Engine.pack_key(rn, mn), tn)  # both slots hold a batch
            acc_a, n_acc = a.enqueue_packed_end()
            acc_b = b.enqueue_packed(h, pkg.Engine.pack_key(rating, mode), ts)
            assert np.array_equal(acc_a, acc_b) and n_acc == int((acc_b == 1).sum())
            ia, ca = a.enqueue_rejects(); ib, cb = b.enqueue_rejects()
            assert np.array_equal(np.sort(ia), np.sort(ib)) and len(ia) == (50 if k else 0)
            la, ma, sa, _ = a.tick_packed()
            lb, mb, sb, _ = b.tick_packed()
            assert np.array_equal(la, lb) and np.array_equal(ma, mb) and np.array_equal(sa, sb)
            assert len(la) > 0
        assert np.array_equal(a.pool_read()["id"], b.pool_read()["id"])


def test_packed_async_results_do_not_block_the_next_tick(pkg):
    """async_results + mm_tick_packed: a tick's kernels run while the previous tick's host copies are still in flight
    (two device buffer sets); the previous tick's host arrays are complete once the next mm_tick_packed returns.
    Host arrays alternate, as a consumer that reads tick k while tick k+1 runs would."""
    import torch
    n, steps = 250_007, 5
    cfg = pkg.synth.make_config(n_groups=32, order=RATING, capacity=2 * n, active_capacity=(steps + 1) * n)
    cfg.flags |= pkg.abi.MM_F_DENSE_IDS
    lob_h = [torch.empty(n, dtype=torch.int64).pin_memory() for _ in range(2)]
    mem_h = [torch.empty(2 * n, dtype=torch.int32).pin_memory() for _ in range(2)]
    with pkg.Engine(cfg) as a, pkg.Engine(cfg) as b:
        a.set_option("async_results", 1)
        prev = None
        for k in range(steps):
            _, rating, mode, ts = make_pool(pkg, 900 + k, n, oor=0.0)
            handles = np.arange(n, dtype=np.uint32) + np.uint32(k * n)
            key = pkg.Engine.pack_key(rating, mode)
            assert (a.enqueue_packed(handles, key, ts) == 1).all() and (b.enqueue_packed(handles, key, ts) == 1).all()
            st = a.tick_raw(lob_h[k & 1].data_ptr(), n, mem_h[k & 1].data_ptr(), 2 * n, packed=True)
            if prev is not None:  # tick k has returned: the host arrays of tick k-1 are complete
                pst, plob, pmem = prev
                assert np.array_equal(lob_h[(k - 1) & 1].numpy()[:pst.n_lobbies].view(plob.dtype), plob)
                assert np.array_equal(mem_h[(k - 1) & 1].numpy()[:pst.n_matched].view(np.uint32), pmem)
            lob, mem, _, stb = b.tick_packed(want_emit_seq=False)
            assert (st.n_lobbies, st.n_matched) == (stb.n_lobbies, stb.n_matched) and st.n_lobbies > 0
            prev = (st, lob, mem)
        a.results_wait()
        pst, plob, pmem = prev
        assert np.array_equal(lob_h[(steps - 1) & 1].numpy()[:pst.n_lobbies].view(plob.dtype), plob)
        assert np.array_equal(mem_h[(steps - 1) & 1].numpy()[:pst.n_matched].view(np.uint32), pmem)


# ---- rating-group shards on real engines: K engines, one per rank's group range, merged ----------------
@pytest.mark.parametrize("K,workload", [(2, None), (4, None), (8, "config3_10m_g32_5v5")])
@pytest.mark.parametrize("order", [RATING, ARRIVAL])
def test_group_sharded_gpu_engines_merge_to_the_single_engine(pkg, oracle, K, workload, order):
    """shard.route deals the players to K engines by rating group (generic/worker.ex:46-69), every engine ticks on
    its own, shard.merge_results == the single engine over the whole pool == the oracle.  K = 8 on BASELINE
    configs[3] (one 10 M pool, 4 groups per engine)."""
    import importlib
    shard = importlib.import_module("microservice-matchmaking_b200.shard")
    if workload:
        w = pkg.synth.WORKLOADS[workload]
        n, G = w["n"], w["n_groups"]
        cfg = pkg.synth.make_config(n_groups=G, order=order, capacity=n)
        ids, rating, mode, ts = pkg.synth.gen_pool(1, n, mode=w["mode"])
    else:
        n, G = 400_003, 32
        cfg = pkg.synth.make_config(n_groups=G, order=order, capacity=n)
        ids, rating, mode, ts = make_pool(pkg, 31 * K, n, bell=True, oor=0.0)
    owner = shard.route(cfg, rating, K)
    assert (owner >= 0).all()
    per_rank, resid = [], []
    for r in range(K):
        mine = owner == r
        cfg_r = pkg.synth.make_config(n_groups=G, order=order, capacity=max(int(mine.sum()), 1))
        with pkg.Engine(cfg_r) as eng:
            assert eng.enqueue(ids[mine], rating[mine], mode[mine], ts[mine]).all()
            lob, mem, seq, st = eng.tick()
            per_rank.append((lob, mem, None))
            resid.append(eng.pool_read()["id"])
            assert set(np.unique(lob["group"])) <= set(shard.groups_of_rank(r, G, K).tolist())
    mlob, mmem = shard.merge_results(cfg, per_rank)[:2]
    with pkg.Engine(cfg) as eng:
        assert eng.enqueue(ids, rating, mode, ts).all()
        lob1, mem1, _, st1 = eng.tick()
        resid1 = eng.pool_read()["id"]
    assert np.array_equal(mlob, lob1) and np.array_equal(mmem, mem1)
    assert np.array_equal(np.sort(np.concatenate(resid)), np.sort(resid1))
    if n <= 1_000_000:
        ref = oracle.run_closed_form(cfg, ids, rating, mode)
        assert np.array_equal(mlob, ref.lobbies) and np.array_equal(mmem, ref.member_ids)
    else:
        lm, lg, members, res = oracle.closed_form_numpy(cfg, ids, rating, mode)
        assert np.array_equal(members, mmem) and np.array_equal(lg, mlob["group"])
