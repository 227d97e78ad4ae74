#!/bin/bash
# compute-sanitizer over the kernels every step runs (fan-out incl. span, index, single-pass plan, TMA gather, async
# import, commit).  Output: gpurun_out/sanitizer_*.txt; the summary lines go to profiles/r2_sanitizer.txt by hand.
out=gpurun_out; mkdir -p $out
T="tests/test_gpu_parity.py tests/test_gpu_xshard.py tests/test_gpu_n3.py -k "not latency_server""
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 $( [ $tool = synccheck ] && echo --num-cuda-barriers 262144 ) python -m pytest $T -m gpu -q -x -p no:cacheprovider > $out/sanitizer_$tool.txt 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $out/sanitizer_$tool.txt | tr '\n' ' ')"
done
