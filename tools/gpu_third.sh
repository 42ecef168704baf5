#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_engine_gpu.py::test_config3_ten_million_5v5 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest.log
timeout 600 python -m pytest tests -m gpu -q -k "config3" > gpurun_out/pytest_10m.log 2>&1; echo "pytest10m rc=$?"; tail -3 gpurun_out/pytest_10m.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_engine_gpu.py -q -x -k "4097-1-3 or few_distinct" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_engine_gpu.py -q -x -k "4097-1-3 or multi_tick or degenerate" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
timeout 600 python tools/exp_place.py config3_10m_g32_5v5 1 > gpurun_out/exp_rating.log 2>&1; tail -8 gpurun_out/exp_rating.log
timeout 600 python tools/exp_place.py config3_10m_g32_5v5 0 > gpurun_out/exp_arrival.log 2>&1; tail -8 gpurun_out/exp_arrival.log
timeout 600 python tools/exp_place.py config2_1m_g8_1v1 1 > gpurun_out/exp_1m.log 2>&1; tail -8 gpurun_out/exp_1m.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_place|k_hist|k_colscan|k_epilogue" -s 4 -c 4 -o gpurun_out/prof_tick -f python tools/one_tick.py > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches5.csv python tools/one_tick.py > gpurun_out/ncu_l.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
