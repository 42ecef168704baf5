#!/bin/bash
# full single-GPU evidence run (round 2): parity, smoke, bench lines, ncu launch list + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem --format=csv,noheader
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 3 --workload config2_1m_g8_1v1 --no-cpu-baseline --stream-seconds 0 > gpurun_out/bench_1m.log 2>&1; tail -1 gpurun_out/bench_1m.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --order arrival --no-cpu-baseline --no-e2e > gpurun_out/bench_arrival.log 2>&1; tail -1 gpurun_out/bench_arrival.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --two-modes --no-cpu-baseline --no-e2e > gpurun_out/bench_twomodes.log 2>&1; tail -1 gpurun_out/bench_twomodes.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --tick-impl 0 --no-cpu-baseline --no-e2e > gpurun_out/bench_split.log 2>&1; tail -1 gpurun_out/bench_split.log | cut -c1-200
for W in 0 20; do
timeout 300 python bench.py --steps 10 --warmup 3 --max-spread $W --no-cpu-baseline --no-e2e > gpurun_out/bench_s1_w$W.log 2>&1; tail -1 gpurun_out/bench_s1_w$W.log | cut -c1-200
done
timeout 300 python tools/exp_ingest.py > gpurun_out/ingest.txt 2>&1; tail -2 gpurun_out/ingest.txt
timeout 300 python tools/exp_small_tick.py > gpurun_out/small_tick.txt 2>&1; grep impl=1 gpurun_out/small_tick.txt
# launch list of the fused tick with the clocks the profiler saw (reconciles ncu durations with the CUDA-event clock)
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,gpc__cycles_elapsed.avg.per_second,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tick" -s 3 -c 1 -o gpurun_out/prof_tick -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_place|k_hist|k_colscan|k_epilogue" -s 12 -c 4 -o gpurun_out/prof_split -f python bench.py --steps 2 --warmup 3 --tick-impl 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full2.log 2>&1; echo "ncu split rc=$?"
