"""Device side of the ingest alone (inputs already in HBM): mm_enqueue_device of 10 M players, hashed and dense active sets."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("microservice-matchmaking_b200")
n = 10_000_000
for dense in ((1,) if os.environ.get('MM_INGEST_ONCE') else (0, 1)):
    cfg, mi = pkg.synth.workload_config("config3_10m_g32_5v5", 1, n + 65536)
    if dense:
        cfg.flags |= pkg.abi.MM_F_DENSE_IDS; cfg.active_capacity = 4 * n
    ids, rating, mode, ts = pkg.synth.gen_pool(1, n, mode=mi)
    if dense: ids = np.arange(n, dtype=np.uint64)
    d = [torch.from_numpy(x.view(np.int64) if x.dtype == np.uint64 else x).cuda() for x in (ids, rating, mode)]
    with pkg.Engine(cfg) as eng:
        ts_ = []
        for it in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            acc = eng.enqueue_device(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr())
            torch.cuda.synchronize(); ts_.append((time.perf_counter() - t0) * 1e3)
            assert acc == n
            eng.tick_device(); eng.remove(ids)
        print(f"dense={dense}: mm_enqueue_device 10 M players: {min(ts_[1:]):.2f} ms (runs {['%.2f' % x for x in ts_]})")
