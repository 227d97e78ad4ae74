#!/bin/bash
# short GPU check: full GPU suite + smoke + the broadcast-at-scale probe (each under its own timeout)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench_extra.py bcast > gpurun_out/bcast.jsonl 2> gpurun_out/bcast.err; cut -c1-1200 gpurun_out/bcast.jsonl; tail -2 gpurun_out/bcast.err | cut -c1-300
