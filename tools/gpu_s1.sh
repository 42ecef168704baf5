#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rating_window.py -m gpu -q -x > gpurun_out/pytest_s1.log 2>&1; echo "pytest s1 rc=$?"; tail -3 gpurun_out/pytest_s1.log
for W in 0 2 20; do
timeout 300 python bench.py --steps 10 --warmup 3 --max-spread $W --no-cpu-baseline --no-e2e > gpurun_out/bench_s1_w$W.log 2>&1
done
timeout 300 python tools/stream_bench.py rate=1e6 seconds=2 dt_ms=1 max_spread=5 > gpurun_out/stream_s1.jsonl 2>&1; cat gpurun_out/stream_s1.jsonl | cut -c1-600
