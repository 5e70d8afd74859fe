#!/bin/bash
# session K: derived (prefix-slice) searches, build streams, LFA kernel breakdown, narrow wgrad
mkdir -p gpurun_out/r02k
python -m pytest tests/test_gpu_knn.py tests/test_gpu_pass.py tests/test_gpu_train.py tests/test_gpu_backproject.py -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02k/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02k/pytest.log
for nb in 0 2 4; do
FFB6D_BUILD_STREAMS=$nb python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mlp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('build streams $nb', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['digest_ok'], d['reference_digest_ok'], d['gpu_launches'])"
done
python tools/lfa_times.py 32 > gpurun_out/r02k/lfa_times.txt 2>&1; tail -22 gpurun_out/r02k/lfa_times.txt
python tools/train_bench.py --config 3 --steps 3 --warmup 2 --profile 2> gpurun_out/r02k/train_profile.txt | cut -c1-200; sed -n 4,14p gpurun_out/r02k/train_profile.txt
