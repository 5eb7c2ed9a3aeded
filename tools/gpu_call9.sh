#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c9_gpu_tests.log 2>&1; echo "pytest all rc=$?"; grep -n "passed\|failed" gpurun_out/c9_gpu_tests.log | tail -2
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/c9_memcheck_state.log python -m pytest tests/test_gpu_state.py -m gpu -x -q > gpurun_out/c9_state.log 2>&1; echo "sanitizer state rc=$?"; tail -3 gpurun_out/c9_state.log; grep -c "Invalid" gpurun_out/c9_memcheck_state.log; head -40 gpurun_out/c9_memcheck_state.log
