#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --workload config2_1m_g8_1v1 --no-cpu-baseline --no-e2e > gpurun_out/bench_1m.log 2>&1; tail -1 gpurun_out/bench_1m.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --order arrival --no-cpu-baseline --no-e2e > gpurun_out/bench_arrival.log 2>&1; tail -1 gpurun_out/bench_arrival.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --two-modes --no-cpu-baseline --no-e2e > gpurun_out/bench_twomodes.log 2>&1; tail -1 gpurun_out/bench_twomodes.log | cut -c1-200
