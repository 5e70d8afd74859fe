#!/bin/bash
mkdir -p gpurun_out/r02o
timeout 600 python -m pytest tests/test_gpu_pose.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/r02o/pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r02o/pytest.log
timeout 300 python tools/pose_times.py 2>&1 | tee gpurun_out/r02o/pose_times.txt
