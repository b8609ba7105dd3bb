#!/bin/bash
# One GPU-box round trip: grouped GPU tests, smoke, default bench (both arms), phase profile, kernel probe.
mkdir -p gpurun_out
export OPENBLAS_NUM_THREADS=${OPENBLAS_NUM_THREADS:-8}
TEST_TIMEOUT=600 tools/gpu_tests.sh tests/test_gpu_kernels.py "test" tests/test_gpu_solvers.py "test"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout 300 python tools/phase_profile.py > gpurun_out/phase.log 2>&1; tail -n 4 gpurun_out/phase.log
timeout 600 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; echo "bench_kernels rc=$?"
