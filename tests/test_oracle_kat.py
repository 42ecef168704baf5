"""Known-answer tests of the CPU oracle, derived from rules the reference's own code
pins.  The reference ships no golden vectors for the search path
(parity unpinned), so these KATs are the anchor: each cites the reference lines."""
import numpy as np
import pytest

abi = None
synth = None


@pytest.fixture(autouse=True)
def _mods(pkg):
    global abi, synth
    abi, synth = pkg.abi, pkg.synth


def ref_cfg(order=0, **kw):
    return synth.make_config(groups=synth.REFERENCE_GROUPS, order=order, **kw)


# ---- bucketing: generic/worker.ex:26-27,46-53 + config/config.exs:27-36 -----------------
@pytest.mark.parametrize("rating,name", [
    (0, "bronze"), (1499, "bronze"), (1500, "silver"), (1999, "silver"), (2000, "gold"),
    (2999, "platinum"), (3000, "diamond"), (3999, "master"), (4000, "grandmaster"),
    (5000, "grandmaster"), (5001, "diamond"), (-1, "diamond"), (1499.5, "diamond"),
])
def test_bucketing_reference_defaults(oracle, rating, name):
    cfg = ref_cfg()
    g = oracle.find_rating_group(cfg, rating)
    assert synth.REFERENCE_GROUP_NAMES[g] == name


def test_default_group_index(oracle):
    L = oracle.lib()
    assert L.orc_default_group_index(7) == 4  # div(7, 2) + 1 -> "diamond"
    assert L.orc_default_group_index(1) == -1  # Enum.at out of range -> nil -> MatchError
    assert L.orc_default_group_index(2) == -1
    assert L.orc_default_group_index(3) == 2
    assert ref_cfg().default_group == 4


def test_first_match_in_list_order(oracle):
    # overlapping ranges: Enum.find returns the FIRST tuple that matches
    cfg = synth.make_config(groups=[(0, 100), (50, 200), (0, 1000)], default_group=-1)
    assert oracle.find_rating_group(cfg, 75) == 0
    assert oracle.find_rating_group(cfg, 150) == 1
    assert oracle.find_rating_group(cfg, 500) == 2
    assert oracle.find_rating_group(cfg, 1001) == -1


def test_out_of_range_without_default_is_error(oracle):
    cfg = synth.make_config(n_groups=1)  # G=1 -> default undefined
    assert cfg.default_group == -1
    with pytest.raises(ValueError):
        oracle.run_literal(cfg, [1], [6000], [0])


# ---- one search worker, serialized: search/worker.ex:291-324 -----------------------------
def test_empty_pool(oracle):
    r = oracle.run_literal(ref_cfg(), [], [], [])
    assert r.n_lobbies == 0 and r.n_matched == 0 and r.n_residual == 0


def test_1v1_pairs_in_arrival_order(oracle):
    cfg = ref_cfg(order=abi.MM_ORDER_ARRIVAL)
    ids = np.arange(100, 107, dtype=np.uint64)
    rating = [100, 1600, 200, 1700, 300, 400, 4500]  # bronze: 100,102,104,105  silver: 101,103  gm: 106
    r = oracle.run_literal(cfg, ids, rating, np.zeros(7, np.uint8))
    assert r.n_lobbies == 3
    got = [tuple(r.member_ids[h["first_member"]:h["first_member"] + h["n_members"]]) for h in r.lobbies]
    # canonical order: (mode, group, emission)
    assert got == [(100, 102), (104, 105), (101, 103)]
    assert list(r.lobbies["group"]) == [0, 0, 1]
    # emission order of the serialized loop: lobby completes when its last member arrives
    assert list(r.emit_seq) == [2, 5, 3]
    assert list(r.emission_rank) == [0, 2, 1]
    assert list(r.residual_ids) == [106]


def test_residual_counts(oracle):
    # n players in one partition => floor(n/L) lobbies and n mod L still queued
    cfg = ref_cfg()
    for n in (0, 1, 9, 10, 11, 19, 20, 25):
        ids = np.arange(1, n + 1, dtype=np.uint64)
        r = oracle.run_literal(cfg, ids, np.full(n, 2100), np.ones(n, np.uint8))  # 5v5, gold
        assert r.n_lobbies == n // 10 and r.n_residual == n % 10
        assert list(r.residual_ids) == list(ids[(n // 10) * 10:])


def test_lobby_size_and_team_major(oracle):
    # game-lobby/worker.ex:37-39: sum of team lengths = L; S0 fills "team 1" first
    cfg = ref_cfg()
    ids = np.arange(1, 11, dtype=np.uint64)
    r = oracle.run_literal(cfg, ids, np.full(10, 3100), np.ones(10, np.uint8))
    assert r.n_lobbies == 1 and r.lobbies[0]["n_members"] == 10
    assert list(r.member_ids) == list(range(1, 11))  # team 1 = first five joiners


def test_independence_of_group_and_mode(oracle):
    # lobby_state.ex:72-79: state is selected by table (=group) and game_mode
    cfg = ref_cfg()
    ids = np.arange(1, 5, dtype=np.uint64)
    r = oracle.run_literal(cfg, ids, [100, 1600, 100, 100], [0, 0, 1, 0])
    assert r.n_lobbies == 1
    assert sorted(r.member_ids) == [1, 4]  # same group AND same mode only
    assert sorted(r.residual_ids) == [2, 3]


def test_leaver_is_filtered_not_matched(oracle):
    # search/worker.ex:267-280,312-321: a member that left is filtered; a lobby whose
    # membership changed is NOT emitted, the shrunken state is saved
    cfg = ref_cfg()
    ids = np.array([1, 2, 3, 4, 5], np.uint64)
    alive = np.array([1, 0, 1, 1, 1], np.uint8)
    r = oracle.run_literal(cfg, ids, np.full(5, 100), np.zeros(5, np.uint8), alive=alive)
    got = [tuple(r.member_ids[h["first_member"]:h["first_member"] + 2]) for h in r.lobbies]
    assert got == [(1, 3), (4, 5)]
    assert r.n_dead == 1 and r.n_residual == 0
    assert list(r.emit_seq) == [2, 4]


def test_leaver_last_joiner(oracle):
    cfg = ref_cfg()
    r = oracle.run_literal(cfg, [1, 2], [100, 100], [0, 0], alive=[1, 0])
    assert r.n_lobbies == 0 and list(r.residual_ids) == [1] and r.n_dead == 1


def test_rating_order_feed(oracle):
    # ORDER_RATING: (mode, clamp(rating), enqueue order)
    cfg = ref_cfg(order=abi.MM_ORDER_RATING)
    ids = np.array([10, 11, 12, 13, 14], np.uint64)
    r = oracle.run_literal(cfg, ids, [900, 100, 500, 100, 1400], np.zeros(5, np.uint8))
    got = [tuple(r.member_ids[h["first_member"]:h["first_member"] + 2]) for h in r.lobbies]
    assert got == [(11, 13), (12, 10)]  # 100,100 (tie by enqueue order) then 500,900
    assert list(r.residual_ids) == [14]


def test_rating_order_default_group_quirk(oracle):
    # out-of-range ratings land in "diamond" (index 4): -1 sorts first, 5001 last
    cfg = ref_cfg(order=abi.MM_ORDER_RATING)
    ids = np.array([1, 2, 3, 4], np.uint64)
    r = oracle.run_literal(cfg, ids, [5001, 3200, -1, 3100], np.zeros(4, np.uint8))
    assert list(r.lobbies["group"]) == [4, 4]
    got = [tuple(r.member_ids[h["first_member"]:h["first_member"] + 2]) for h in r.lobbies]
    assert got == [(3, 4), (2, 1)]


# ---- literal loop == closed form (the property the GPU engine is held to) ------------
@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_literal_equals_closed_form(oracle, order, seed):
    rng = np.random.default_rng(seed)
    n = 5000
    cfg = synth.make_config(groups=synth.REFERENCE_GROUPS, order=order)
    ids, rating, _, _ = synth.gen_pool(seed, n)
    rating = rating.copy()
    rating[rng.integers(0, n, 50)] = rng.integers(-50, 5100, 50)  # some out-of-range
    mode = rng.integers(0, 2, n).astype(np.uint8)
    alive = (rng.random(n) > 0.05).astype(np.uint8)
    a = oracle.run_literal(cfg, ids, rating, mode, alive)
    b = oracle.run_closed_form(cfg, ids, rating, mode, alive)
    assert a.n_lobbies == b.n_lobbies and a.n_matched == b.n_matched and a.n_requeued == 0
    assert np.array_equal(a.lobbies, b.lobbies)
    assert np.array_equal(a.member_ids, b.member_ids)
    assert np.array_equal(a.emit_seq, b.emit_seq)
    assert np.array_equal(a.residual_ids, b.residual_ids)
    lm, lg, mem, res = oracle.closed_form_numpy(cfg, ids, rating, mode, alive)
    assert np.array_equal(mem, a.member_ids) and np.array_equal(res, a.residual_ids)
    assert np.array_equal(lm, a.lobbies["mode"]) and np.array_equal(lg, a.lobbies["group"])
    if order == 0:  # arrival: emission order == sort by completing member's arrival
        assert np.array_equal(np.argsort(a.emit_seq, kind="stable"), np.argsort(a.emission_rank, kind="stable"))


def test_config1_1k_one_group_1v1(oracle):
    # BASELINE.json configs[0]: 1k synthetic players, 1 rating group, 1v1
    cfg = synth.make_config(n_groups=1, order=abi.MM_ORDER_ARRIVAL)
    ids, rating, mode, _ = synth.gen_pool(1, 1000)
    r = oracle.run_literal(cfg, ids, rating, mode)
    assert r.n_lobbies == 500 and r.n_residual == 0
    assert np.array_equal(r.member_ids, ids)  # arrival order, pairs (0,1),(2,3),...


def test_timed_legs_agree(oracle):
    cfg = synth.make_config(n_groups=8, order=abi.MM_ORDER_RATING)
    ids, rating, mode, _ = synth.gen_pool(2, 20000)
    s1, l1 = oracle.time_literal(cfg, ids, rating, mode, 1)
    s4, l4 = oracle.time_literal(cfg, ids, rating, mode, 4)
    ref = oracle.run_closed_form(cfg, ids, rating, mode)
    assert l1 == l4 == ref.n_lobbies and s1 > 0 and s4 > 0


# ---- generator twins ----------------------------------------------------------------------
@pytest.mark.parametrize("bell", [False, True])
def test_generator_c_equals_numpy(oracle, bell):
    a = oracle.gen_pool_c(3, 4096, first=12345, bell=bell, mode=1)
    b = synth.gen_pool(3, 4096, first=12345, bell=bell, mode=1)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert len(np.unique(a[0])) == 4096
    assert a[1].min() >= 0 and a[1].max() <= 5000


def test_equal_width_groups_cover():
    for G in (1, 7, 8, 32):
        los, his = synth.equal_width_groups(G)
        assert los[0] == 0 and his[-1] == 5000
        assert all(his[g] + 1 == los[g + 1] for g in range(G - 1))
