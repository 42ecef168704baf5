#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "random_pool or both_rank or degenerate or few_distinct or leavers" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 600 python tools/exp_place.py config3_10m_g32_5v5 1 > gpurun_out/exp_rating.log 2>&1; cat gpurun_out/exp_rating.log
