#!/bin/bash
# per-kernel ncu captures of the four-launch tick (config3, rating order) + launch list of the fused tick
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_place|k_hist|k_colscan|k_epilogue" -s 12 -c 4 -o gpurun_out/prof_split -f python bench.py --steps 2 --warmup 3 --tick-impl 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full2.log 2>&1; echo "ncu split rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --tick-impl 0 --no-cpu-baseline --no-e2e > gpurun_out/bench_split.log 2>&1; tail -1 gpurun_out/bench_split.log | cut -c1-1200
