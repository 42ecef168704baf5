"""Generates the committed golden fixtures tests/golden/*.npz.

The reference (Elixir) cannot be executed in this environment and ships no vectors for the search path
(parity unpinned — see oracle/mm_oracle.h), so these fixtures are produced by the ORACLE's literal
serialized loop (oracle/mm_oracle.c: orc_run_literal) on seeded inputs.  They pin the oracle (and through
it the GPU engine) against regressions: any change to the oracle, the generator or the engine that alters
a lobby -> player_id assignment shows up as a diff against these files.
    python tests/golden/make_golden.py      # rewrites the fixtures (review the diff before committing)
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("microservice-matchmaking_b200")
oracle = importlib.import_module("oracle.oracle")
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (groups | n_groups, modes, order, n, seed, bell, out_of_range fraction, dead fraction)
    "reference_groups_arrival": (pkg.synth.REFERENCE_GROUPS, pkg.synth.MODES_DEFAULT, 0, 3000, 11, False, 0.02, 0.05),
    "reference_groups_rating": (pkg.synth.REFERENCE_GROUPS, pkg.synth.MODES_DEFAULT, 1, 3000, 12, True, 0.02, 0.05),
    "config1_1k_one_group_1v1": (1, (("1v1", 2, 1),), 0, 1000, 1, False, 0.0, 0.0),
    "g8_rating_three_modes": (8, (("1v1", 2, 1), ("5v5", 2, 5), ("3v3v3", 3, 3)), 1, 5000, 13, False, 0.0, 0.03),
    # EXTENSION (policy S1, rating window — see WINDOW below): produced by orc_run_windowed
    "g8_rating_window6": (8, (("1v1", 2, 1), ("5v5", 2, 5)), 1, 5000, 14, False, 0.0, 0.03),
    "reference_groups_window25": (pkg.synth.REFERENCE_GROUPS, pkg.synth.MODES_DEFAULT, 1, 3000, 15, True, 0.02, 0.05),
}
WINDOW = {"g8_rating_window6": 6, "reference_groups_window25": 25}  # max lobby spread; absent = reference policy S0


def run_oracle(name, cfg, ids, rating, mode, alive):
    if name in WINDOW:
        return oracle.run_windowed(cfg, WINDOW[name], ids, rating, mode, alive)
    return oracle.run_literal(cfg, ids, rating, mode, alive)


def build(name):
    groups, modes, order, n, seed, bell, oor, dead = CASES[name]
    kw = dict(groups=groups) if not isinstance(groups, int) else dict(n_groups=groups)
    cfg = pkg.synth.make_config(modes=modes, order=order, capacity=n, **kw)
    rng = np.random.default_rng(seed)
    ids, rating, _, ts = pkg.synth.gen_pool(seed, n, bell=bell)
    rating = rating.copy()
    k = int(n * oor)
    if k and cfg.default_group >= 0:
        rating[rng.integers(0, n, k)] = rng.integers(-100, 5200, k)
    mode = rng.integers(0, len(modes), n).astype(np.uint8)
    alive = (rng.random(n) >= dead).astype(np.uint8)
    return cfg, ids, rating, mode, alive


if __name__ == "__main__":
    for name in CASES:
        cfg, ids, rating, mode, alive = build(name)
        r = run_oracle(name, cfg, ids, rating, mode, alive)
        rank = r.emission_rank if r.emission_rank is not None else np.zeros(0, np.uint32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), ids=ids, rating=rating, mode=mode, alive=alive,
                            lobbies=r.lobbies, member_ids=r.member_ids, emit_seq=r.emit_seq,
                            emission_rank=rank, residual_ids=r.residual_ids)
        print(name, r.n_lobbies, "lobbies", r.n_matched, "matched", r.n_residual, "residual", r.n_dead, "dead")
