#!/bin/bash
# short GPU check: full GPU suite + smoke + N=1 bench line (each under its own timeout)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --steps 32 --warmup 4 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err | cut -c1-400
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print(round(d["value"]/1e9,3), round(d["ms_per_step"],4), {k:round(x["ms_per_launch"],4) for k,x in d["kernels"].items()}, round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]/1e6,1), "launches", d["gpu_launches"], "p50", d["p50_dequeue_us"])
print("p50 launched", d["configs"]["p50_dequeue"].get("launched_path_p50_us")); print("parity", d.get("parity", {}).get("match"), "ref-python", (d.get("cpu_baseline_reference") or {}).get("value"))
for k, c in (d.get("configs") or {}).items():
    print(" ", k, {x: (round(y, 3) if isinstance(y, float) else y) for x, y in c.items() if x in ("value", "p50_us", "p50_sweep_wall_ms", "k5_kernels_ms_per_sweep", "check", "error", "send_msgs_per_s", "receive_msgs_per_s")}, (c.get("roofline") or {}).get("frac"))
PY
