#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_engine_gpu.py::test_config3_ten_million_5v5 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_engine_gpu.py -q -x -k "fused and 4097" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_engine_gpu.py -q -x -k "fused and (4097 or 300_001 or 2049)" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
for w in "config3_10m_g32_5v5 1" "config3_10m_g32_5v5 0" "config2_1m_g8_1v1 1"; do timeout 600 python tools/exp_place.py $w 2>&1 | tail -3; done | tee gpurun_out/exp_fused.log
timeout 600 python -m pytest tests -m gpu -q -k "config3" > gpurun_out/pytest_10m.log 2>&1; echo "pytest10m rc=$?"; tail -3 gpurun_out/pytest_10m.log
