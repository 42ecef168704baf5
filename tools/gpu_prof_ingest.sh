#!/bin/bash
# ncu capture of the ingest kernels (dense active set, 10 M players, device-resident inputs)
mkdir -p gpurun_out
MM_INGEST_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_enq_" -s 10 -c 4 -o gpurun_out/prof_ingest -f python tools/exp_ingest.py > gpurun_out/ncu_ingest.log 2>&1; echo "ncu ingest rc=$?"
tail -3 gpurun_out/ncu_ingest.log
