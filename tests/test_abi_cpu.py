"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports
every symbol include/mm_engine.h declares; host-only entry points agree with the oracle.
No compute calls here (they need a GPU and live in test_engine_gpu.py)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(pkg):
    import __graft_entry__ as g
    g.build()
    return pkg.load_library()


def test_header_symbols_all_exported(pkg, lib):
    hdr = open(os.path.join(ROOT, "include", "mm_engine.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)  # strip comments
    declared = set(re.findall(r"\b(mm_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no prototypes found"
    assert declared == set(pkg.abi.PROTOTYPES), declared ^ set(pkg.abi.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name), f"libmm_engine.so does not export {name}"


def test_struct_layouts_match_header(pkg):
    abi = pkg.abi
    assert C.sizeof(abi.LobbyHdr) == 8
    assert C.sizeof(abi.TickStats) == 48
    assert C.sizeof(abi.ModeDesc) == 4
    assert C.sizeof(abi.Config) == 4 * 2 + 4 * 64 * 2 + 4 + 4 + 4 * 8 + 4 * 5


def test_default_config_is_the_reference(pkg, lib):
    cfg = pkg.abi.Config()
    lib.mm_config_default(C.byref(cfg))
    assert cfg.n_groups == 7 and cfg.default_group == 4
    assert [(cfg.group_lo[g], cfg.group_hi[g]) for g in range(7)] == list(pkg.synth.REFERENCE_GROUPS)
    assert (cfg.modes[0].teams, cfg.modes[0].team_size) == (2, 1)
    assert (cfg.modes[1].teams, cfg.modes[1].team_size) == (2, 5)


def test_group_of_equals_oracle(pkg, lib, oracle):
    cfg = pkg.synth.make_config(groups=pkg.synth.REFERENCE_GROUPS)
    for r in list(range(-3, 5005)) + [10 ** 6, -10 ** 6]:
        assert lib.mm_group_of(C.byref(cfg), r) == oracle.find_rating_group(cfg, r)
    cfg2 = pkg.synth.make_config(groups=[(0, 100), (50, 200), (0, 1000)], default_group=-1)
    for r in range(-5, 1010):
        assert lib.mm_group_of(C.byref(cfg2), r) == oracle.find_rating_group(cfg2, r)


def test_strerror_and_version(pkg, lib):
    assert lib.mm_abi_version() == pkg.abi.MM_ABI_VERSION
    assert lib.mm_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.mm_strerror(pkg.abi.MM_E_CUDA)


def test_bad_config_rejected_before_any_cuda_call(pkg, lib):
    cfg = pkg.synth.make_config(n_groups=8)
    cfg.n_groups = 0
    h = C.c_void_p()
    assert lib.mm_create(C.byref(cfg), C.byref(h)) == pkg.abi.MM_E_ARG


def test_create_fails_loudly_without_gpu(pkg, lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.EngineError) as ei:
        pkg.Engine(pkg.synth.make_config(n_groups=8))
    assert ei.value.status == pkg.abi.MM_E_CUDA  # no silent CPU path


def test_product_never_touches_the_oracle():
    """The product path may not import / link / execute anything under oracle/."""
    pk = os.path.join(ROOT, "microservice-matchmaking_b200")
    for dp, _dn, fn in os.walk(pk):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liborc" not in txt, f  # never dlopen'ed / linked
                assert not re.search(r"#\s*include\s*[<\"][^>\"]*oracle", txt), f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert not re.search(r"import_module\(\s*[\"']oracle", txt), f
    out = os.popen(f"ldd {pk}/csrc/libmm_engine.so").read()
    assert "liborc" not in out
