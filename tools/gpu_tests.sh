#!/bin/bash
# Runs on the GPU box: GPU tests in small groups (pytest -k pattern), each group in its own process under its own
# timeout so that one hung kernel cannot eat the whole call.  Logs into gpurun_out/.
#   usage: tools/gpu_tests.sh <file> <k-pattern> [<file> <k-pattern> ...]
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=${OPENBLAS_NUM_THREADS:-8}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
: > gpurun_out/summary.txt
while [ $# -ge 2 ]; do
  f=$1; k=$2; shift 2
  b=$(basename "$f" .py)_$(echo "$k" | tr -c 'A-Za-z0-9' '_')
  timeout -k 10 ${TEST_TIMEOUT:-300} python -X faulthandler -m pytest "$f" -v -m gpu -k "$k" --tb=short -p no:cacheprovider \
      -o faulthandler_timeout=120 > gpurun_out/$b.log 2>&1
  rc=$?
  echo "== $f -k '$k' exit $rc : $(tail -n 1 gpurun_out/$b.log)" | tee -a gpurun_out/summary.txt
  if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR|E  )|Timeout|Fatal|File \"" gpurun_out/$b.log | head -n 40; fi
done
