"""Synthetic player pools and configs of SURVEY §8(d) / BASELINE.json.

numpy twin of oracle/mm_oracle.c:orc_gen_pool (tests assert they are bit-identical).
All arithmetic is uint64 with wrap-around, like the C.
"""
import numpy as np

from . import abi

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)
_D = np.uint64(0xD1B54A32D192ED03)


def mix64(z):
    """splitmix64 finalizer (a bijection on u64)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z ^ (z >> np.uint64(30))
        z = z * _M1
        z = z ^ (z >> np.uint64(27))
        z = z * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def gen_pool(seed, n, first=0, bell=False, mode=0):
    """-> (id u64[n], rating i32[n], mode u8[n], enq_ts u32[n]); ids unique by construction."""
    i = np.arange(first, first + n, dtype=np.uint64)
    s = np.uint64(seed)
    with np.errstate(over="ignore"):
        x = mix64(s * _G + i)
        pid = mix64(((s + np.uint64(1)) * _D) ^ i)
    if not bell:
        rating = ((x >> np.uint64(32)) % np.uint64(5001)).astype(np.int32)
    else:
        m = np.uint64(0xFFFF)
        u = (x & m) + ((x >> np.uint64(16)) & m) + ((x >> np.uint64(32)) & m) + ((x >> np.uint64(48)) & m)
        rating = ((u * np.uint64(5000)) // np.uint64(4 * 65535)).astype(np.int32)
    modes = np.full(n, mode, dtype=np.uint8)
    ts = i.astype(np.uint32)
    return pid, rating, modes, ts


def equal_width_groups(n_groups, lo=0, hi=5000):
    """G inclusive integer ranges covering [lo, hi]: lo_g = ceil(W*g/G), hi_g = ceil(W*(g+1)/G)-1."""
    w = hi - lo + 1
    los = [lo + -((-w * g) // n_groups) for g in range(n_groups)]
    his = [lo + -((-w * (g + 1)) // n_groups) - 1 for g in range(n_groups)]
    return los, his


MODES_DEFAULT = (("1v1", 2, 1), ("5v5", 2, 5))


def make_config(n_groups=None, groups=None, modes=MODES_DEFAULT, order=abi.MM_ORDER_RATING,
                capacity=1 << 20, active_capacity=0, device=0, default_group="reference", flags=0):
    """Build an abi.Config.  groups: list of (lo, hi); n_groups: equal-width over [0, 5000].
    default_group "reference" = generic/worker.ex:27's div(len, 2) + 1 (or -1 when out of range)."""
    cfg = abi.Config()
    cfg.abi_version = abi.MM_ABI_VERSION
    if groups is None:
        los, his = equal_width_groups(n_groups if n_groups else 7)
        groups = list(zip(los, his))
    assert 0 < len(groups) <= abi.MM_MAX_GROUPS
    cfg.n_groups = len(groups)
    for g, (lo, hi) in enumerate(groups):
        cfg.group_lo[g] = lo
        cfg.group_hi[g] = hi
    if default_group == "reference":
        idx = len(groups) // 2 + 1
        default_group = idx if idx < len(groups) else -1
    cfg.default_group = default_group
    assert 0 < len(modes) <= abi.MM_MAX_MODES
    cfg.n_modes = len(modes)
    for m, (_name, teams, size) in enumerate(modes):
        cfg.modes[m].teams = teams
        cfg.modes[m].team_size = size
    cfg.order_mode = order
    cfg.capacity = capacity
    cfg.active_capacity = active_capacity
    cfg.device = device
    cfg.flags = flags
    return cfg


REFERENCE_GROUPS = ((0, 1499), (1500, 1999), (2000, 2499), (2500, 2999), (3000, 3499), (3500, 3999), (4000, 5000))
REFERENCE_GROUP_NAMES = ("bronze", "silver", "gold", "platinum", "diamond", "master", "grandmaster")

# BASELINE.json configs (index = position in "configs"); "mode" indexes MODES_DEFAULT
WORKLOADS = {
    "config1_1k_g1_1v1": dict(n=1_000, n_groups=1, mode=0),
    "config2_1m_g8_1v1": dict(n=1_000_000, n_groups=8, mode=0),
    "config3_10m_g32_5v5": dict(n=10_000_000, n_groups=32, mode=1),
}


def workload_config(name, order, capacity, device=0, single_mode=True):
    """(cfg, mode_index_in_cfg) for a BASELINE.json workload.  single_mode=True configures only the
    game mode the workload names ("1v1" or "5v5"), which halves the engine's key domain."""
    w = WORKLOADS[name]
    if single_mode:
        cfg = make_config(n_groups=w["n_groups"], modes=(MODES_DEFAULT[w["mode"]],), order=order, capacity=capacity,
                          device=device)
        return cfg, 0
    return make_config(n_groups=w["n_groups"], order=order, capacity=capacity, device=device), w["mode"]
