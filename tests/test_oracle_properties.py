"""Property tests of the CPU oracle on random configurations (hypothesis): overlapping / gapped rating-group
tables, any default group, several team shapes, leavers.  Four independent restatements must agree:
the literal consume/5 loop, the closed form in C, the closed form in numpy, and (RATING order) the
windowed walk with an unlimited window — plus the conservation laws every tick obeys."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

ARRIVAL, RATING = 0, 1

group_tables = st.lists(st.tuples(st.integers(0, 60), st.integers(0, 25)), min_size=1, max_size=6).map(
    lambda xs: [(lo, lo + w) for lo, w in xs])
mode_tables = st.lists(st.tuples(st.integers(1, 4), st.integers(1, 3)), min_size=1, max_size=3)


@settings(max_examples=120, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(groups=group_tables, modes=mode_tables, order=st.sampled_from([ARRIVAL, RATING]),
       seed=st.integers(0, 2 ** 31), n=st.integers(0, 400), default_pick=st.integers(0, 5), dead=st.floats(0, 0.3))
def test_restatements_agree(pkg, oracle, groups, modes, order, seed, n, default_pick, dead):
    default_group = default_pick % len(groups)
    cfg = pkg.synth.make_config(groups=groups, modes=[(f"m{i}", t, s) for i, (t, s) in enumerate(modes)], order=order,
                                capacity=max(n, 1), default_group=default_group)
    rng = np.random.default_rng(seed)
    ids = pkg.synth.mix64(np.arange(n, dtype=np.uint64) + np.uint64(seed))
    rating = rng.integers(-10, 100, n).astype(np.int32)  # in range, in gaps and out of range (-> default group)
    mode = rng.integers(0, len(modes), n).astype(np.uint8)
    alive = (rng.random(n) >= dead).astype(np.uint8)
    lit = oracle.run_literal(cfg, ids, rating, mode, alive)
    cf = oracle.run_closed_form(cfg, ids, rating, mode, alive)
    assert np.array_equal(lit.lobbies, cf.lobbies) and np.array_equal(lit.member_ids, cf.member_ids)
    assert np.array_equal(lit.residual_ids, cf.residual_ids) and np.array_equal(lit.emit_seq, cf.emit_seq)
    lm, lg, mem, resid = oracle.closed_form_numpy(cfg, ids, rating, mode, alive)
    assert np.array_equal(lm, lit.lobbies["mode"]) and np.array_equal(lg, lit.lobbies["group"])
    assert np.array_equal(mem, lit.member_ids) and np.array_equal(resid, lit.residual_ids)
    if order == RATING:
        w = oracle.run_windowed(cfg, -1, ids, rating, mode, alive)
        assert np.array_equal(w.lobbies, lit.lobbies) and np.array_equal(w.member_ids, lit.member_ids)
        assert np.array_equal(w.residual_ids, lit.residual_ids)
    # conservation: every alive player is matched exactly once or still queued; dead ones are dropped
    assert lit.n_matched + lit.n_residual + lit.n_dead == n and lit.n_dead == int((alive == 0).sum())
    assert len(np.unique(lit.member_ids)) == lit.n_matched
    assert not np.isin(lit.member_ids, lit.residual_ids).any()
    assert not np.isin(ids[alive == 0], np.concatenate([lit.member_ids, lit.residual_ids])).any()
    # lobby shape and independence of (mode, group) (lobby_state.ex:72-79)
    group_of = {int(p): oracle.find_rating_group(cfg, int(r)) for p, r in zip(ids, rating)}
    mode_of = dict(zip(ids.tolist(), mode.tolist()))
    for h in lit.lobbies:
        L = cfg.modes[int(h["mode"])].teams * cfg.modes[int(h["mode"])].team_size
        m = lit.member_ids[h["first_member"]:h["first_member"] + h["n_members"]]
        assert h["n_members"] == L
        assert {mode_of[int(x)] for x in m} == {int(h["mode"])} and {group_of[int(x)] for x in m} == {int(h["group"])}
    # fewer than L players of any partition stay queued (policy S0)
    left = {}
    for x in lit.residual_ids.tolist():
        k = (mode_of[x], group_of[x])
        left[k] = left.get(k, 0) + 1
    for (m, _g), c in left.items():
        assert c < cfg.modes[m].teams * cfg.modes[m].team_size


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 2 ** 31), n=st.integers(0, 300), W=st.integers(0, 12), teams=st.integers(1, 3), size=st.integers(1, 3))
def test_window_properties(pkg, oracle, seed, n, W, teams, size):
    """S1 (extension): lobbies respect the window, nobody is lost, and no window the greedy walk could still fill is
    left behind at the front of a partition."""
    cfg = pkg.synth.make_config(groups=[(0, 30), (31, 60)], modes=[("m", teams, size)], order=RATING, default_group=1)
    rng = np.random.default_rng(seed)
    ids = pkg.synth.mix64(np.arange(n, dtype=np.uint64) + np.uint64(seed))
    rating = rng.integers(-3, 65, n).astype(np.int32)
    mode = np.zeros(n, np.uint8)
    r = oracle.run_windowed(cfg, W, ids, rating, mode)
    L = teams * size
    key = dict(zip(ids.tolist(), np.clip(rating, -1, 61).tolist()))
    assert r.n_matched + r.n_residual == n and r.n_matched == r.n_lobbies * L
    for h in r.lobbies:
        ks = [key[int(x)] for x in r.member_ids[h["first_member"]:h["first_member"] + L]]
        assert ks == sorted(ks) and ks[-1] - ks[0] <= W
    wide = oracle.run_windowed(cfg, 1000, ids, rating, mode)
    lit = oracle.run_literal(cfg, ids, rating, mode)
    assert np.array_equal(wide.member_ids, lit.member_ids)  # a window wider than the domain is S0
    assert r.n_lobbies <= lit.n_lobbies


@pytest.mark.gpu
@settings(max_examples=50, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(groups=group_tables, modes=mode_tables, order=st.sampled_from([ARRIVAL, RATING]), seed=st.integers(0, 2 ** 31),
       n=st.integers(0, 5000), default_pick=st.integers(0, 5), dead=st.floats(0, 0.3), W=st.integers(-1, 10),
       tick_impl=st.sampled_from([0, 1]))
def test_engine_matches_oracle_on_random_configs(pkg, oracle, groups, modes, order, seed, n, default_pick, dead, W, tick_impl):
    """The CUDA tick through the C ABI on the same random configurations (overlapping / gapped group tables,
    several team shapes, leavers, both orders, both policies): bit-exact against the oracle."""
    cfg = pkg.synth.make_config(groups=groups, modes=[(f"m{i}", t, s) for i, (t, s) in enumerate(modes)], order=order,
                                capacity=max(n, 1), default_group=default_pick % len(groups))
    rng = np.random.default_rng(seed)
    ids = pkg.synth.mix64(np.arange(n, dtype=np.uint64) + np.uint64(seed))
    rating = rng.integers(-10, 100, n).astype(np.int32)
    mode = rng.integers(0, len(modes), n).astype(np.uint8)
    alive = (rng.random(n) >= dead).astype(np.uint8)
    if order == ARRIVAL:
        W = -1
    ref = oracle.run_windowed(cfg, W, ids, rating, mode, alive) if W >= 0 else oracle.run_literal(cfg, ids, rating, mode, alive)
    with pkg.Engine(cfg) as eng:
        eng.set_option("tick_impl", tick_impl)
        eng.set_option("max_spread", W)
        assert eng.enqueue(ids, rating, mode).all()
        eng.remove(ids[alive == 0])
        lob, mem, seq, st_ = eng.tick()
        assert (st_.n_lobbies, st_.n_matched, st_.n_residual, st_.n_dead) == (ref.n_lobbies, ref.n_matched, ref.n_residual, ref.n_dead)
        assert np.array_equal(lob, ref.lobbies) and np.array_equal(mem, ref.member_ids) and np.array_equal(seq, ref.emit_seq)
        assert np.array_equal(eng.pool_read()["id"], ref.residual_ids)
