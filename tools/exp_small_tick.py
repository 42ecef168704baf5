"""Phase times of small ticks (fused vs four launches): python tools/exp_small_tick.py"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
pkg = importlib.import_module("microservice-matchmaking_b200")
for n, G in ((1000, 32), (10_000, 32), (100_000, 32), (1_000_000, 8)):
    modes = (("5v5", 2, 5),) if G == 32 else (("1v1", 2, 1),)
    cfg = pkg.synth.make_config(n_groups=G, modes=modes, order=1, capacity=n + 1000)
    ids, rating, _, ts = pkg.synth.gen_pool(5, n)
    mode = np.zeros(n, np.uint8)
    for impl in (1, 0):
        with pkg.Engine(cfg) as eng:
            eng.set_option("tick_impl", impl)
            eng.enqueue(ids, rating, mode, ts); eng.snapshot()
            rows = []
            for it in range(12):
                eng.restore()
                st = eng.tick_device()
                if it >= 2: rows.append((st.device_us, st.hist_us, st.scan_us, st.place_us, st.epilogue_us))
            m = np.median(np.array(rows), axis=0)
            print(f"n={n:8d} G={G:2d} impl={impl} launches={st.n_launches} device {m[0]:6.1f} us | hist {m[1]:5.1f} scan {m[2]:5.1f} place {m[3]:5.1f} epi {m[4]:5.1f}")
