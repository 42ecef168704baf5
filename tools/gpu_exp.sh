#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log
for w in "config3_10m_g32_5v5 1 single" "config3_10m_g32_5v5 0 single" "config2_1m_g8_1v1 1 single"; do echo "== $w"; timeout 600 python tools/exp_place.py $w 2>&1 | tail -2; done | tee gpurun_out/exp_bincol.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phase_us'], d['roofline']['frac'], d['e2e'])"
