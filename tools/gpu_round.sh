#!/bin/bash
# One GPU-box round trip: grouped GPU tests, smoke, default bench (both arms), phase profile, kernel probe.
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=${OPENBLAS_NUM_THREADS:-8}
TEST_TIMEOUT=600 tools/gpu_tests.sh tests/test_gpu_kernels.py "test" tests/test_gpu_solvers.py "test" tests/test_gpu_multi.py "test"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
for t in ${EXTRA_TOOLS}; do timeout 300 python tools/$t > gpurun_out/${t%.py}.log 2>&1; echo "$t rc=$?"; tail -n 6 gpurun_out/${t%.py}.log; done
