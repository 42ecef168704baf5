#!/bin/bash
# rating-window extension: parity first, then the regression suite and bench lines
mkdir -p gpurun_out
python tools/one_tick.py config3_10m_g32_5v5 1
timeout 900 python -m pytest tests/test_rating_window.py -m gpu -q -x > gpurun_out/pytest_s1.log 2>&1; echo "pytest s1 rc=$?"; tail -5 gpurun_out/pytest_s1.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_rating_window.py > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --tick-impl 0 --no-cpu-baseline --no-e2e > gpurun_out/bench_split.log 2>&1; tail -1 gpurun_out/bench_split.log | cut -c1-300
