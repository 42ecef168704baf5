#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_engine_gpu.py::test_config3_ten_million_5v5 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
for w in "config3_10m_g32_5v5 1 single" "config3_10m_g32_5v5 1"; do echo "== $w"; timeout 600 python tools/exp_place.py $w 2>&1 | tail -6; done | tee gpurun_out/exp_hist.log
