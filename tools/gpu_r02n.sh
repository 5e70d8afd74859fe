#!/bin/bash
mkdir -p gpurun_out/r02n
python -m pytest tests/test_gpu_fusion_mlp.py tests/test_gpu_fusion_stage.py tests/test_gpu_model.py tests/test_gpu_train.py -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02n/pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r02n/pytest.log
echo "--- MT2"; python tools/mlp_bench.py 2>&1 | tee gpurun_out/r02n/mlp_mt2.txt | cut -c1-200
echo "--- MT1"; FFB6D_MLP_NO_MT2=1 python tools/mlp_bench.py 2>&1 | tee gpurun_out/r02n/mlp_mt1.txt | cut -c1-200
