#!/bin/bash
# round-2 call C: device witness assignment + pi circuit after the short-circuit products
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_assign.py tests/test_gpu_pi.py -m gpu -q -x > $O/c_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -12 $O/c_gpu_tests.log
timeout 600 python bench.py --workload assign --steps 20 > $O/c_wl_assign.json 2> $O/c_wl_assign.err; echo "assign rc=$?"; tail -c 1500 $O/c_wl_assign.json; tail -3 $O/c_wl_assign.err
timeout 600 python bench.py --workload pi --steps 20 > $O/c_wl_pi.json 2> $O/c_wl_pi.err; echo "pi rc=$?"; tail -c 700 $O/c_wl_pi.json; tail -3 $O/c_wl_pi.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_assign|k_seg|k_check' -c 60 --csv --log-file $O/c_assign_launches.csv python bench.py --workload assign --steps 3 > /dev/null 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py $O/c_assign_launches.csv 2
