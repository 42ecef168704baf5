#!/bin/bash
# first GPU call: parity tests, smoke, sanitizer, bench, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -5 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_engine_gpu.py::test_config3_ten_million_5v5 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest.log
timeout 600 python -m pytest tests -m gpu -q -k "config3" > gpurun_out/pytest_10m.log 2>&1; echo "pytest10m rc=$?"
tail -8 gpurun_out/pytest_10m.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_engine_gpu.py -q -x -k "kat_leaver or dedupe or 4097 or capacity or multi_tick" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -6 gpurun_out/memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_engine_gpu.py -q -x -k "4097 and 1-3" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
tail -6 gpurun_out/racecheck.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/bench.log
timeout 300 python bench.py --steps 10 --warmup 3 --workload config2_1m_g8_1v1 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1
tail -1 gpurun_out/bench_1m.log
timeout 300 python bench.py --steps 5 --warmup 3 --order arrival --no-cpu-baseline --no-e2e > gpurun_out/bench_arrival.log 2>&1
tail -1 gpurun_out/bench_arrival.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
tail -1 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1; echo "ncu rc=$?"
