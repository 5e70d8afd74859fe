#!/bin/bash
mkdir -p gpurun_out/r02g
python -m pytest tests/test_gpu_model.py tests/test_gpu_backproject.py -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02g/pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r02g/pytest.log; cat gpurun_out/model_grad_errors.json
