#!/bin/bash
mkdir -p gpurun_out/r02m
python -m pytest tests/test_gpu_lfa.py tests/test_gpu_gather.py tests/test_gpu_train.py -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02m/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02m/pytest.log
python tools/lfa_times.py 32 > gpurun_out/r02m/lfa_times.txt 2>&1; sed -n 3,10p gpurun_out/r02m/lfa_times.txt
python tools/train_bench.py --config 3 --steps 3 --warmup 2 --profile 2> gpurun_out/r02m/train_profile.txt | cut -c1-200; sed -n 3,12p gpurun_out/r02m/train_profile.txt
