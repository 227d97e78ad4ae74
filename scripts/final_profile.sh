#!/bin/bash
# Round-end evidence on ONE GPU: full GPU suite, smoke, N=1 bench line, ncu launch list, ncu full captures.
# Every command runs under its own timeout; outputs land in gpurun_out/final/.
out=gpurun_out/final; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/pytest_gpu.txt
SDB_FANOUT_VARIANT=3 timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/pytest_gpu_variant3.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1
timeout 300 python bench.py --steps 64 --warmup 4 > $out/bench_n1.json 2> $out/bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file $out/launches.csv \
    python bench.py --steps 4 --warmup 3 --cpu-budget 0 > $out/launch_bench.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none \
    -k 'regex:k_group_fanout_warp|k_recv_gather|k_pull_index|k_recv_select' -s 10 -c 5 -o $out/prof_n1 -f \
    python bench.py --steps 4 --warmup 3 --cpu-budget 0 > $out/ncu_n1.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_group_fanout_span -s 2 -c 1 -o $out/prof_span_w8 -f \
    python scripts/proxy_import8.py > $out/ncu_span.log 2>&1
cat $out/pytest_gpu.txt $out/pytest_gpu_variant3.txt $out/smoke.txt
tail -c 600 $out/bench_n1.json
