#!/bin/bash
# ncu evidence for the N=1 step on ONE GPU: launch list (cold-cache, serialised: compare SHARES) + one full capture of
# the step's kernels.  Outputs land in gpurun_out/prof/ (scratch); summaries are copied to profiles/ by hand.
out=gpurun_out/prof; mkdir -p $out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 16 --csv --log-file $out/launches.csv \
    python bench.py --steps 4 --warmup 3 --cpu-budget 0 --no-extras > $out/launch_bench.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none \
    -k 'regex:k_group_fanout|k_recv_gather_tma|k_pull_index_group|k_recv_plan' -s 8 -c 4 -o $out/prof_n1 -f \
    python bench.py --steps 4 --warmup 3 --cpu-budget 0 --no-extras > $out/ncu_n1.log 2>&1
ls -la $out; tail -3 $out/ncu_n1.log
