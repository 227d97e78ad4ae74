#!/bin/bash
# K1 (point-to-point enqueue) roofline line + one ncu capture of its kernels
mkdir -p gpurun_out/prof
python bench_extra.py k1 2>/dev/null | tail -1 | tee gpurun_out/k1.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_enqueue_p2p|k_commit" -s 6 -c 2 -o gpurun_out/prof/k1 -f python bench_extra.py k1 > gpurun_out/prof/k1_ncu.log 2>&1
tail -2 gpurun_out/prof/k1_ncu.log
