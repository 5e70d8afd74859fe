#!/bin/bash
# session J: fused LFA kernel parity + timing, training-step wins, KNN occupancy-cap experiment
mkdir -p gpurun_out/r02j
python -m pytest tests/test_gpu_lfa.py tests/test_gpu_train.py tests/test_gpu_model.py -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02j/pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02j/pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02j/bench.err > gpurun_out/r02j/bench.json; python -c "
import json
d=json.load(open('gpurun_out/r02j/bench.json'))
print('pass', d['ms_per_step'], 'lfa', d['lfa_blocks'])"
for n in 3 4; do
FFB6D_KNN_MAX_CTAS=$n python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mlp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('knn max ctas $n', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['digest_ok'], round(d['compute']['knn_ms_per_step'],3))"
done
python tools/train_bench.py --config 3 --steps 3 --warmup 2 --profile 2> gpurun_out/r02j/train_profile.txt | cut -c1-300; head -16 gpurun_out/r02j/train_profile.txt
