#!/bin/bash
# Runs on the GPU box: every GPU test file under its own timeout, logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/smi.txt 2>&1
for f in "$@"; do
  b=$(basename "$f" .py)
  timeout -k 10 ${TEST_TIMEOUT:-900} python -m pytest "$f" -q -m gpu -x --tb=short -p no:cacheprovider > gpurun_out/$b.log 2>&1
  echo "== $f exit $?" | tee -a gpurun_out/summary.txt
  tail -n 30 gpurun_out/$b.log
done
