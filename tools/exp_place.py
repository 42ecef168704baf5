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
cfg = pkg.synth.make_config(n_groups=w["n_groups"], order=order, capacity=n + 65536)
ids, rating, mode, ts = pkg.synth.gen_pool(1, n, mode=w["mode"])
eng = pkg.Engine(cfg)
assert eng.enqueue(ids, rating, mode, ts).all()
eng.snapshot()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(label, reps=6, **opts):
    for k, v in opts.items():
        eng.set_option(k, v)
    dev, pl = [], []
    for i in range(reps):
        eng.restore(); flush.fill_(1); torch.cuda.synchronize()
        st = eng.tick_device()
        if i >= 2: dev.append(st.device_us); pl.append(st.place_us)
    print(json.dumps({"variant": label, "workload": name, "order": order, "device_us": round(float(np.mean(dev)), 1),
                      "place_us": round(float(np.mean(pl)), 1), "place_min": round(float(np.min(pl)), 1), "lobbies": st.n_lobbies}), flush=True)
run("impl3 auto (block512 x2 rows/SM)")
run("impl3 block1024 rows1 stages4", rows_per_sm=1, block=1024, place2_stages=4)
run("impl3 block512 rows1 stages4", block=512)
run("impl3 block512 rows2 stages2", rows_per_sm=2, place2_stages=2)
run("impl3 block512 rows2 stages2 dense=0", dense=0)
run("impl1 legacy", rank_impl=1, rows_per_sm=1, block=1024)
