#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --log-file gpurun_out/c6_memcheck_packed.log python -m pytest "tests/test_gpu_packed.py::test_every_golden_family_with_packed_storage" -x -q > gpurun_out/c6_packed.log 2>&1; echo "sanitizer packed rc=$?"; tail -3 gpurun_out/c6_packed.log; grep -c "Invalid\|ERROR" gpurun_out/c6_memcheck_packed.log; head -60 gpurun_out/c6_memcheck_packed.log
