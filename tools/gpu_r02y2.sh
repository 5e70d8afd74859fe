#!/bin/bash
O=gpurun_out/r02y2; mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_knn.py tests/test_gpu_gather.py tests/test_gpu_grid.py tests/test_gpu_pass.py tests/test_gpu_fusion_mlp.py tests/test_gpu_fusion_stage.py tests/test_gpu_model.py -m gpu -q --timeout 1400 -p no:cacheprovider > $O/memcheck_rest.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error:" $O/memcheck_rest.log | cut -c1-250 | head -12
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_knn.py tests/test_gpu_pose.py tests/test_gpu_backproject.py -m gpu -q --timeout 800 -p no:cacheprovider -k "golden or organised or sweep or matches_reference or masks or sampl or index" > $O/initcheck.log 2>&1
echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Uninitialized|Error:" $O/initcheck.log | cut -c1-250 | head -12
