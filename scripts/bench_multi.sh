#!/bin/bash
# usage: scripts/bench_multi.sh N [steps]  - N-GPU bench line under torchrun, summary on stdout
N=$1; K=${2:-32}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $K --warmup 4 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N", round(d["value"]/1e9,3), "G msg/s", round(d["ms_per_step"],4), "ms", {k:round(x["ms_per_launch"],4) for k,x in d["kernels"].items()}, d["phases_ms_rank0"], "e2e", round(d["e2e"]["value"]/1e6,1), "M/s", round(d["e2e"]["ms_per_step"],2), "ms parity", d["parity"]["match"])
PY
