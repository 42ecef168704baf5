#!/bin/bash
# full single-GPU evidence run: parity, smoke, bench lines, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 300 python bench.py --steps 20 --warmup 3 --workload config2_1m_g8_1v1 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1; tail -1 gpurun_out/bench_1m.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --order arrival --no-cpu-baseline --no-e2e > gpurun_out/bench_arrival.log 2>&1; tail -1 gpurun_out/bench_arrival.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --two-modes --no-cpu-baseline --no-e2e > gpurun_out/bench_twomodes.log 2>&1; tail -1 gpurun_out/bench_twomodes.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --tick-impl 0 --no-cpu-baseline --no-e2e > gpurun_out/bench_split.log 2>&1; tail -1 gpurun_out/bench_split.log | cut -c1-300
for W in 0 2 20; do
timeout 300 python bench.py --steps 10 --warmup 3 --max-spread $W --no-cpu-baseline > gpurun_out/bench_s1_w$W.log 2>&1; tail -1 gpurun_out/bench_s1_w$W.log | cut -c1-300
done
timeout 300 python tools/stream_bench.py rate=1e6 seconds=2 dt_ms=1,5 > gpurun_out/stream.jsonl 2>&1; cat gpurun_out/stream.jsonl | cut -c1-500
timeout 300 python tools/stream_bench.py rate=1e6 seconds=2 dt_ms=1 max_spread=5 > gpurun_out/stream_s1.jsonl 2>&1; cat gpurun_out/stream_s1.jsonl | cut -c1-500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tick" -s 3 -c 1 -o gpurun_out/prof_tick -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_place2|k_hist3|k_colscan|k_epilogue" -s 12 -c 4 -o gpurun_out/prof_split -f python bench.py --steps 2 --warmup 3 --tick-impl 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full2.log 2>&1; echo "ncu split rc=$?"
