#!/bin/bash
mkdir -p gpurun_out
for w in "config3_10m_g32_5v5 0 single" "config3_10m_g32_5v5 1 single"; do echo "== $w"; timeout 600 python tools/exp_place.py $w 2>&1 | tail -6; done | tee gpurun_out/exp_arrival_decomp.log
