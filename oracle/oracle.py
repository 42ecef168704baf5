"""ctypes wrapper of the CPU oracle (oracle/liborc.so) + a numpy closed form.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
PARITY UNPINNED: see oracle/mm_oracle.h.
"""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
abi = importlib.import_module("microservice-matchmaking_b200.abi")


class OrcResult(C.Structure):
    _fields_ = [
        ("n_lobbies", C.c_uint32),
        ("n_matched", C.c_uint64),
        ("n_residual", C.c_uint32),
        ("n_dead", C.c_uint32),
        ("n_requeued", C.c_uint32),
        ("lobbies", C.POINTER(abi.LobbyHdr)),
        ("member_ids", C.POINTER(C.c_uint64)),
        ("emit_seq", C.POINTER(C.c_uint32)),
        ("emission_rank", C.POINTER(C.c_uint32)),
        ("residual_ids", C.POINTER(C.c_uint64)),
    ]


def build(force=False):
    """Compile oracle/liborc.so with gcc (build the checker; not the product)."""
    so = os.path.join(_HERE, "liborc.so")
    src = [os.path.join(_HERE, "mm_oracle.c"), os.path.join(_HERE, "mm_oracle.h"),
           os.path.join(_ROOT, "include", "mm_engine.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        vp = C.c_void_p
        L.orc_find_rating_group.restype = C.c_int
        L.orc_find_rating_group.argtypes = [C.POINTER(abi.Config), C.c_double]
        L.orc_default_group_index.restype = C.c_int
        L.orc_default_group_index.argtypes = [C.c_uint32]
        L.orc_required_slots.restype = C.c_uint32
        L.orc_required_slots.argtypes = [vp, C.c_uint32]
        for name in ("orc_run_literal", "orc_run_closed_form"):
            fn = getattr(L, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(abi.Config), C.c_uint32, C.c_uint32, vp, vp, vp, vp, C.POINTER(OrcResult)]
        L.orc_run_windowed.restype = C.c_int
        L.orc_run_windowed.argtypes = [C.POINTER(abi.Config), C.c_int32, C.c_uint32, vp, vp, vp, vp, C.POINTER(OrcResult)]
        L.orc_result_free.restype = None
        L.orc_result_free.argtypes = [C.POINTER(OrcResult)]
        L.orc_time_literal.restype = C.c_double
        L.orc_time_literal.argtypes = [C.POINTER(abi.Config), C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint32,
                                       C.POINTER(C.c_uint32)]
        L.orc_session_new.restype = vp
        L.orc_session_new.argtypes = [C.POINTER(abi.Config)]
        L.orc_session_free.restype = None
        L.orc_session_free.argtypes = [vp]
        L.orc_session_feed.restype = C.c_int
        L.orc_session_feed.argtypes = [vp, C.c_uint32, vp, vp, vp, vp]
        L.orc_session_remove.restype = C.c_int
        L.orc_session_remove.argtypes = [vp, C.c_uint32, vp, C.POINTER(C.c_uint32)]
        L.orc_session_take.restype = C.c_int
        L.orc_session_take.argtypes = [vp, C.POINTER(OrcResult), C.POINTER(C.POINTER(C.c_uint8))]
        L.orc_free.restype = None
        L.orc_free.argtypes = [vp]
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint64]
        L.orc_gen_pool.restype = None
        L.orc_gen_pool.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint8, vp, vp, vp, vp]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Result:
    """numpy view of an oracle run: canonical (mode, group, emission) lobby order."""

    def __init__(self, r, with_rank):
        nl, nm = r.n_lobbies, r.n_matched
        hdr = np.ctypeslib.as_array(C.cast(r.lobbies, C.POINTER(C.c_uint8)), shape=(max(nl, 1) * 8,))[: nl * 8]
        hdr = hdr.copy().view(np.dtype([("first_member", "<u4"), ("n_members", "<u2"), ("mode", "u1"), ("group", "u1")]))
        self.lobbies = hdr
        self.member_ids = np.ctypeslib.as_array(r.member_ids, shape=(max(nm, 1),))[:nm].copy()
        self.emit_seq = np.ctypeslib.as_array(r.emit_seq, shape=(max(nl, 1),))[:nl].copy()
        self.emission_rank = (np.ctypeslib.as_array(r.emission_rank, shape=(max(nl, 1),))[:nl].copy()
                              if with_rank and r.emission_rank else None)
        nr = r.n_residual
        self.residual_ids = np.ctypeslib.as_array(r.residual_ids, shape=(max(nr, 1),))[:nr].copy()
        self.n_lobbies, self.n_matched, self.n_residual = nl, nm, nr
        self.n_dead, self.n_requeued = r.n_dead, r.n_requeued


def _run(fn, cfg, order, ids, rating, mode, alive, with_rank):
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    rating = np.ascontiguousarray(rating, dtype=np.int32)
    mode = np.ascontiguousarray(mode, dtype=np.uint8)
    if alive is not None:
        alive = np.ascontiguousarray(alive, dtype=np.uint8)
    r = OrcResult()
    rc = fn(C.byref(cfg), order, len(ids), _ptr(ids), _ptr(rating), _ptr(mode), _ptr(alive), C.byref(r))
    if rc != 0:
        raise ValueError(f"oracle returned {rc}")
    try:
        return Result(r, with_rank)
    finally:
        lib().orc_result_free(C.byref(r))


def run_literal(cfg, ids, rating, mode, alive=None, order=None):
    """The serialized consume/5 loop (search/worker.ex:291-324) over the queued set."""
    return _run(lib().orc_run_literal, cfg, cfg.order_mode if order is None else order, ids, rating, mode, alive, True)


def run_closed_form(cfg, ids, rating, mode, alive=None, order=None):
    return _run(lib().orc_run_closed_form, cfg, cfg.order_mode if order is None else order, ids, rating, mode, alive,
                False)


def run_windowed(cfg, max_spread, ids, rating, mode, alive=None):
    """EXTENSION: policy S1 (a lobby spans at most max_spread rating points), RATING order; < 0 = unlimited."""
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    rating = np.ascontiguousarray(rating, dtype=np.int32)
    mode = np.ascontiguousarray(mode, dtype=np.uint8)
    if alive is not None:
        alive = np.ascontiguousarray(alive, dtype=np.uint8)
    r = OrcResult()
    rc = lib().orc_run_windowed(C.byref(cfg), int(max_spread), len(ids), _ptr(ids), _ptr(rating), _ptr(mode), _ptr(alive),
                                C.byref(r))
    if rc != 0:
        raise ValueError(f"oracle returned {rc}")
    try:
        return Result(r, True)
    finally:
        lib().orc_result_free(C.byref(r))


def time_literal(cfg, ids, rating, mode, n_threads=1, order=None):
    """-> (seconds, lobbies) for the literal loop on the host cores."""
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    rating = np.ascontiguousarray(rating, dtype=np.int32)
    mode = np.ascontiguousarray(mode, dtype=np.uint8)
    nl = C.c_uint32(0)
    s = lib().orc_time_literal(C.byref(cfg), cfg.order_mode if order is None else order, len(ids), _ptr(ids),
                               _ptr(rating), _ptr(mode), n_threads, C.byref(nl))
    if s < 0:
        raise ValueError("oracle timing failed")
    return s, nl.value


def find_rating_group(cfg, rating):
    return lib().orc_find_rating_group(C.byref(cfg), float(rating))


def gen_pool_c(seed, n, first=0, bell=False, mode=0):
    ids = np.empty(n, np.uint64)
    rating = np.empty(n, np.int32)
    modes = np.empty(n, np.uint8)
    ts = np.empty(n, np.uint32)
    lib().orc_gen_pool(seed, first, n, int(bell), mode, _ptr(ids), _ptr(rating), _ptr(modes), _ptr(ts))
    return ids, rating, modes, ts


def closed_form_numpy(cfg, ids, rating, mode, alive=None, order=None):
    """Vectorised closed form (for pools too large for the literal loop in a test):
    -> (lobby_mode, lobby_group, members[n_lobbies_total], residual_ids) with members
    lobby-major in canonical order.  Group rule = generic/worker.ex:46-53."""
    order = cfg.order_mode if order is None else order
    ids = np.asarray(ids, np.uint64)
    rating = np.asarray(rating, np.int64)
    mode = np.asarray(mode, np.int64)
    n = len(ids)
    G = cfg.n_groups
    grp = np.full(n, cfg.default_group, np.int64)
    unset = np.ones(n, bool)
    for g in range(G):
        hit = unset & (rating >= cfg.group_lo[g]) & (rating <= cfg.group_hi[g])
        grp[hit] = g
        unset &= ~hit
    keep = np.ones(n, bool) if alive is None else np.asarray(alive, bool)
    idx = np.nonzero(keep)[0]
    if order == abi.MM_ORDER_RATING:
        rmin = min(cfg.group_lo[g] for g in range(G))
        rmax = max(cfg.group_hi[g] for g in range(G))
        ck = np.clip(rating[idx], rmin - 1, rmax + 1)
        feed = idx[np.lexsort((idx, ck, mode[idx]))]
    else:
        feed = idx
    seg = mode[feed] * G + grp[feed]
    o = np.argsort(seg, kind="stable")
    part = feed[o]
    seg_s = seg[o]
    members, lm, lg, resid = [], [], [], []
    bounds = np.searchsorted(seg_s, np.arange(cfg.n_modes * G + 1))
    for s in range(cfg.n_modes * G):
        a, b = bounds[s], bounds[s + 1]
        m = s // G
        L = cfg.modes[m].teams * cfg.modes[m].team_size
        nl = (b - a) // L
        members.append(ids[part[a:a + nl * L]])
        lm.append(np.full(nl, m, np.uint8))
        lg.append(np.full(nl, s % G, np.uint8))
        resid.append(part[a + nl * L:b])
    resid = np.sort(np.concatenate(resid)) if resid else np.zeros(0, np.int64)
    return (np.concatenate(lm), np.concatenate(lg), np.concatenate(members), ids[resid])


class Session:
    """The reference's state kept ACROSS requests (orc_session_*): LobbyState rows survive between batches, the
    active set is mutated by feed / remove, requests are consumed one at a time in arrival order."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.h = lib().orc_session_new(C.byref(cfg))
        if not self.h:
            raise ValueError("orc_session_new failed")

    def close(self):
        if self.h:
            lib().orc_session_free(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def feed(self, ids, rating, mode):
        ids = np.ascontiguousarray(ids, np.uint64); rating = np.ascontiguousarray(rating, np.int32)
        mode = np.ascontiguousarray(mode, np.uint8)
        acc = np.empty(len(ids), np.uint8)
        rc = lib().orc_session_feed(self.h, len(ids), _ptr(ids), _ptr(rating), _ptr(mode), _ptr(acc))
        if rc:
            raise ValueError(f"orc_session_feed returned {rc}")
        return acc

    def remove(self, ids):
        ids = np.ascontiguousarray(ids, np.uint64)
        k = C.c_uint32(0)
        lib().orc_session_remove(self.h, len(ids), _ptr(ids), C.byref(k))
        return k.value

    def take(self):
        """-> (Result of the lobbies emitted since the last take, hole flags u8[n_lobbies])."""
        r = OrcResult()
        hp = C.POINTER(C.c_uint8)()
        rc = lib().orc_session_take(self.h, C.byref(r), C.byref(hp))
        if rc:
            raise ValueError(f"orc_session_take returned {rc}")
        try:
            res = Result(r, True)
            hole = np.ctypeslib.as_array(hp, shape=(max(res.n_lobbies, 1),))[:res.n_lobbies].copy()
            return res, hole
        finally:
            lib().orc_free(hp)
            lib().orc_result_free(C.byref(r))
