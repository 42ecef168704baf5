#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['phase_us'], d['e2e'])"
timeout 300 python bench.py --steps 10 --warmup 3 --workload config2_1m_g8_1v1 --no-cpu-baseline > gpurun_out/bench_1m.log 2>&1; python -c "
import json; d=json.loads(open('gpurun_out/bench_1m.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['phase_us'], d['e2e'])"
