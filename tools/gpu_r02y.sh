#!/bin/bash
O=gpurun_out/r02y; mkdir -p $O
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_pose.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "matches_reference or masks" > $O/racecheck_pose.log 2>&1
echo "racecheck pose rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" $O/racecheck_pose.log | head -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_lfa.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "ffb6d_ds0 or ffb6d_ds2 or ffb6d_ds3 or blk_8_16" > $O/racecheck_lfa.log 2>&1
echo "racecheck lfa rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|Race reported" $O/racecheck_lfa.log | cut -c1-250 | head -12
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_knn.py tests/test_gpu_gather.py -m gpu -q -x --timeout 800 -p no:cacheprovider -k "golden or organised or large_shapes" > $O/racecheck_knn.log 2>&1
echo "racecheck knn/gather rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|Race reported" $O/racecheck_knn.log | cut -c1-250 | head -12
