#!/bin/bash
# gpurun_out/ of tools/gpu_round.sh -> the tracked evidence files profiles/rNN_* (usage: tools/collect_profiles.sh r02)
set -e
R=${1:-r02}
for f in bench bench_ref bench_1m bench_arrival bench_twomodes bench_split bench_s1_w0 bench_s1_w20 bench_n2 bench_n8; do
  [ -s gpurun_out/$f.log ] && tail -1 gpurun_out/$f.log | python -c "import json,sys; print(json.dumps(json.loads(sys.stdin.read()), indent=1))" > profiles/${R}_$f.json || true
done
grep -v "^==PROF==" gpurun_out/launches.csv > profiles/${R}_launches_fused_config3_rating.csv
python tools/ncu_summary.py gpurun_out/prof_tick.ncu-rep > profiles/${R}_ncu_full_k_tick_summary.csv
python tools/ncu_summary.py gpurun_out/prof_split.ncu-rep > profiles/${R}_ncu_full_split_kernels_summary.csv
[ -s gpurun_out/prof_ingest.ncu-rep ] && python tools/ncu_summary.py gpurun_out/prof_ingest.ncu-rep > profiles/${R}_ncu_full_ingest_kernels_summary.csv
cp gpurun_out/ingest.txt profiles/${R}_ingest_device_side.txt
grep "impl=" gpurun_out/small_tick.txt > profiles/${R}_small_ticks.txt
