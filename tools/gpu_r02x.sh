#!/bin/bash
# compute-sanitizer passes over the round-2 kernels (memcheck on a test subset, racecheck on the shared-memory heavy ones)
O=gpurun_out/r02x; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_pass.py -m gpu -q --timeout 600 -p no:cacheprovider -k wrap > $O/pytest_wrap.log 2>&1; echo "wrap rc=$?"; tail -2 $O/pytest_wrap.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_pose.py tests/test_gpu_lfa.py tests/test_gpu_train.py tests/test_gpu_backproject.py -m gpu -q -x --timeout 1100 -p no:cacheprovider > $O/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" $O/memcheck.log | head -12
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_pose.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "matches_reference or masks" > $O/racecheck_pose.log 2>&1
echo "racecheck pose rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/racecheck_pose.log | head -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_lfa.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "fused" > $O/racecheck_lfa.log 2>&1
echo "racecheck lfa rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" $O/racecheck_lfa.log | head -8
