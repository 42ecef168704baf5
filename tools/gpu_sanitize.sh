#!/bin/bash
mkdir -p gpurun_out
K='(random_pool or window_parity or dedupe or leaver or multi_tick or few_bins or sparse_pool or schedule or worker_scenario or packed or persistent or boundary or kat) and not 600011 and not 2000003 and not 300001'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$K" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck.log | tail -3
K2='(window_parity and (2049 or 70001)) or (random_pool and 4097) or multi_tick_accumulation or packed_rejects'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$K2" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/racecheck.log | tail -3
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$K2" > gpurun_out/synccheck.log 2>&1; echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/synccheck.log | tail -3
