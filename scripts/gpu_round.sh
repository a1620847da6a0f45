#!/usr/bin/env bash
# Run under gpurun: each test file in its own process (a trapped kernel kills its CUDA context), logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in "$@"; do
  name=$(basename "$f" .py)
  echo "=== $f" 
  timeout 900 python -m pytest "$f" -x -q -m gpu --timeout 300 -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "exit $? : $(tail -n 3 gpurun_out/$name.log | tr '\n' ' ')"
done
