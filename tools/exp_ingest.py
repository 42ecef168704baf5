"""Ingest decomposition: PCIe copy alone, device-side ingest alone, mm_enqueue (pipelined)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("microservice-matchmaking_b200")
n = 10_000_000
cfg, m = pkg.synth.workload_config("config3_10m_g32_5v5", 1, n + 65536)
ids, rating, mode, ts = pkg.synth.gen_pool(1, n, mode=m)
pin = lambda a: torch.from_numpy(a).pin_memory()
h = [pin(ids.view(np.int64)), pin(rating), pin(mode), pin(ts.view(np.int32))]
d = [torch.empty_like(x, device="cuda") for x in h]
acc = torch.empty(n, dtype=torch.uint8).pin_memory()
def t_copy():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for a, b in zip(d, h): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
print("H2D 170 MB alone ms:", [round(t_copy(), 3) for _ in range(4)])
eng = pkg.Engine(cfg)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.enqueue_raw(n, h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h[3].data_ptr(), acc.data_ptr())
    t1 = time.perf_counter()
    eng.tick_device(); eng.remove(ids)
    print("mm_enqueue (host buffers) ms:", round((t1 - t0) * 1e3, 3))
lib, hh = eng.lib, eng.h
import ctypes as C
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nacc = C.c_uint32(0)
    rc = lib.mm_enqueue_device(hh, n, C.c_void_p(d[0].data_ptr()), C.c_void_p(d[1].data_ptr()), C.c_void_p(d[2].data_ptr()),
                               C.c_void_p(d[3].data_ptr()), None, C.byref(nacc))
    t1 = time.perf_counter()
    assert rc == 0 and nacc.value == n, (rc, nacc.value)
    eng.tick_device(); eng.remove(ids)
    print("mm_enqueue_device (kernels only) ms:", round((t1 - t0) * 1e3, 3))
