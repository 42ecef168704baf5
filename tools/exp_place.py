"""Timing experiments for the placement kernel (GPU).  Prints device/place microseconds
for the L2-hint / persisting-window / debug variants on one workload."""
import importlib, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
pkg = importlib.import_module("microservice-matchmaking_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "config3_10m_g32_5v5"
order = int(sys.argv[2]) if len(sys.argv) > 2 else 1
w = pkg.synth.WORKLOADS[name]
n = w["n"]
single = len(sys.argv) > 3 and sys.argv[3] == "single"
if single:  # only the workload's own game mode in the mode table
    m = pkg.synth.MODES_DEFAULT[w["mode"]]
    cfg = pkg.synth.make_config(n_groups=w["n_groups"], order=order, capacity=n + 65536, modes=(m,))
    ids, rating, mode, ts = pkg.synth.gen_pool(1, n, mode=0)
else:
    cfg = pkg.synth.make_config(n_groups=w["n_groups"], order=order, capacity=n + 65536)
    ids, rating, mode, ts = pkg.synth.gen_pool(1, n, mode=w["mode"])
eng = pkg.Engine(cfg)
assert eng.enqueue(ids, rating, mode, ts).all()
eng.snapshot()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(label, reps=6, **opts):
    for k, v in opts.items():
        eng.set_option(k, v)
    dev, pl, hs, sc, ep = [], [], [], [], []
    for i in range(reps):
        eng.restore(); flush.fill_(1); torch.cuda.synchronize()
        st = eng.tick_device()
        if i >= 2: dev.append(st.device_us); pl.append(st.place_us); hs.append(st.hist_us); sc.append(st.scan_us); ep.append(st.epilogue_us)
    print(json.dumps({"variant": label, "workload": name, "order": order, "device_us": round(float(np.mean(dev)), 1),
                      "place_us": round(float(np.mean(pl)), 1), "place_min": round(float(np.min(pl)), 1), "hist_us": round(float(np.mean(hs)), 1), "scan_us": round(float(np.mean(sc)), 1), "epi_us": round(float(np.mean(ep)), 1), "tail_us": st.reserved / 100.0, "lobbies": st.n_lobbies}), flush=True)
run("fused", tick_impl=1)
run("split", tick_impl=0)
