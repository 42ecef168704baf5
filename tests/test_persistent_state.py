"""The reference keeps LobbyState ACROSS requests (models/lobby_state.ex:61-131, search/worker.ex:312-321); the
batched tick re-derives the partial lobby from the players left resident.  These tests pin the relation:

  * without leavers the two are identical, tick by tick (same lobbies, same member order, same emission batch);
  * with leavers the lobby MEMBERSHIP and the batch in which a lobby comes out are still identical; only the
    member order (= the team split) of a lobby that saw a member leave while it was being filled may differ —
    the reference fills the hole in that team first, the tick orders by enqueue.  DESIGN.md §2 documents it.

CPU half: known-answer tests of the persistent oracle from the reference's own rules.  GPU half: the same
enqueue / remove / tick schedule through the C ABI."""
import numpy as np
import pytest

ARRIVAL = 0


def lobby_lists(lob, mem):
    return [tuple(int(x) for x in mem[h["first_member"]:h["first_member"] + h["n_members"]]) for h in lob]


def cfg_2v2(pkg, cap=64):
    return pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS, modes=(("2v2", 2, 2),), order=ARRIVAL, capacity=cap)


def test_persistent_hole_is_filled_first(pkg, oracle):
    """search/worker.ex:312-321: a member of the saved partial lobby left -> the lobby is not emitted on the request
    that would have filled it, the shrunken state is saved, the next joiner takes the hole in team 1."""
    A, B, Cc, D, E = 1, 2, 3, 4, 5
    with oracle.Session(cfg_2v2(pkg)) as s:
        assert list(s.feed([A, B, Cc], [100] * 3, [0] * 3)) == [1, 1, 1]
        assert s.remove([B]) == 1
        assert list(s.feed([D], [100], [0])) == [1]
        r, hole = s.take()
        assert r.n_lobbies == 0 and sorted(r.residual_ids) == [A, Cc, D]      # filled && changed -> saved, not emitted
        s.feed([E], [100], [0])
        r, hole = s.take()
        assert lobby_lists(r.lobbies, r.member_ids) == [(A, E, Cc, D)] and list(hole) == [1]  # E filled team 1's hole
        assert list(r.emit_seq) == [4]


def test_persistent_without_leavers_equals_fresh_literal_runs(pkg, oracle):
    """No leavers: the persistent session and a fresh literal run per batch over the carried residual agree exactly."""
    cfg = pkg.synth.make_config(n_groups=8, order=ARRIVAL, capacity=100_000)
    rng = np.random.default_rng(3)
    queued = [np.zeros(0, np.uint64), np.zeros(0, np.int32), np.zeros(0, np.uint8)]
    with oracle.Session(cfg) as s:
        first = 0
        for step in range(5):
            n = int(rng.integers(1, 4000))
            ids, rating, _, _ = pkg.synth.gen_pool(5, n, first=first); first += n
            mode = rng.integers(0, 2, n).astype(np.uint8)
            assert s.feed(ids, rating, mode).all()
            queued = [np.concatenate([q, x]) for q, x in zip(queued, (ids, rating, mode))]
            ref = oracle.run_literal(cfg, *queued)
            r, hole = s.take()
            assert not hole.any()
            assert np.array_equal(r.lobbies, ref.lobbies) and np.array_equal(r.member_ids, ref.member_ids)
            assert sorted(r.residual_ids) == sorted(ref.residual_ids)
            keep = np.isin(queued[0], ref.residual_ids)
            queued = [q[keep] for q in queued]


def test_duplicate_and_invalid_requests(pkg, oracle):
    with oracle.Session(cfg_2v2(pkg)) as s:
        assert list(s.feed([1, 1, 2], [100, 100, 9999], [0, 0, 0])) == [1, 0, 1]   # 9999 -> default group "diamond"
        assert list(s.feed([3], [100], [5])) == [2]                                  # unknown game mode
        assert s.remove([1, 77]) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_tick_equals_persistent_reference_up_to_the_team_split(pkg, oracle, seed):
    """Random schedule of enqueue batches, leavers (queued and matched players) and ticks, arrival order, several
    groups and team shapes: per tick the same lobbies come out (as member sets, in the same canonical order), and
    every lobby that did not see a mid-lobby leaver has the identical member order (= identical teams)."""
    modes = (("1v1", 2, 1), ("2v2", 2, 2), ("5v5", 2, 5), ("3x3", 3, 3))
    cfg = pkg.synth.make_config(n_groups=4, modes=modes, order=ARRIVAL, capacity=200_000)
    rng = np.random.default_rng(seed)
    holes_seen = exact = 0
    with pkg.Engine(cfg) as eng, oracle.Session(cfg) as s:
        first = 0
        queued = np.zeros(0, np.uint64)
        for step in range(12):
            n = int(rng.integers(1, 3000))
            ids, rating, _, _ = pkg.synth.gen_pool(11, n, first=first); first += n
            mode = rng.integers(0, len(modes), n).astype(np.uint8)
            assert (eng.enqueue(ids, rating, mode) == s.feed(ids, rating, mode)).all()
            queued = np.concatenate([queued, ids])
            lob, mem, seq, st = eng.tick()
            r, hole = s.take()
            assert st.n_lobbies == r.n_lobbies and st.n_matched == r.n_matched
            assert np.array_equal(lob, r.lobbies)  # same (mode, group, size) sequence and member offsets
            got, want = lobby_lists(lob, mem), lobby_lists(r.lobbies, r.member_ids)
            for g_, w_, h_ in zip(got, want, hole):
                assert sorted(g_) == sorted(w_)
                if not h_:
                    assert g_ == w_
                    exact += 1
                else:
                    holes_seen += 1
            matched = np.isin(queued, mem)
            queued = queued[~matched]
            assert sorted(eng.pool_read()["id"]) == sorted(r.residual_ids) == sorted(queued)
            # leavers: some still queued (mid-lobby holes), some unknown ids; the lobby stage removes matched ones
            if len(queued):
                gone = rng.choice(queued, size=min(len(queued), int(rng.integers(0, 6))), replace=False)
                both = np.concatenate([gone, mem[:50], np.array([10 ** 15], np.uint64)])
                assert eng.remove(both) == s.remove(both)
                queued = queued[~np.isin(queued, gone)]
    assert exact > 100 and holes_seen > 0  # the schedule exercised both kinds
