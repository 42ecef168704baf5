#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "arrival_small_key or fused or random_pool or multi_tick or leavers" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --order arrival --no-cpu-baseline --no-e2e > gpurun_out/bench_arrival.log 2>&1; tail -1 gpurun_out/bench_arrival.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('arrival', d['ms_per_step'], d['phase_us'], d['roofline']['frac'])"
timeout 300 python bench.py --steps 10 --warmup 3 --order arrival --two-modes --no-cpu-baseline --no-e2e > gpurun_out/bench_arrival2.log 2>&1; tail -1 gpurun_out/bench_arrival2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('arrival two modes', d['ms_per_step'], d['phase_us'])"
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_engine_gpu.py -q -x -k "arrival_small_key and 2-" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
