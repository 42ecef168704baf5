"""Two ticks of one workload (for ncu captures): python tools/one_tick.py [workload] [order] [opt=val ...]"""
import importlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
pkg = importlib.import_module("microservice-matchmaking_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "config3_10m_g32_5v5"
order = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = pkg.synth.WORKLOADS[name]
n = w["n"]
cfg = pkg.synth.make_config(n_groups=w["n_groups"], order=order, capacity=n + 65536)
ids, rating, mode, ts = pkg.synth.gen_pool(1, n, mode=w["mode"])
eng = pkg.Engine(cfg)
for kv in sys.argv[3:]:
    k, v = kv.split("="); eng.set_option(k, int(v))
assert eng.enqueue(ids, rating, mode, ts).all()
eng.snapshot()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(2):
    eng.restore(); flush.fill_(1); torch.cuda.synchronize()
    st = eng.tick_device()
    print(i, st.n_lobbies, "resid", st.n_residual, "dev", round(st.device_us, 1), "hist", round(st.hist_us, 1), "scan", round(st.scan_us, 1),
          "place", round(st.place_us, 1), "epi", round(st.epilogue_us, 1), "dbg", st.reserved / 100.0)
