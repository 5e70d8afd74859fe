#!/bin/bash
mkdir -p gpurun_out/r02q
timeout 900 python -m pytest tests/test_gpu_gather.py tests/test_gpu_pass.py tests/test_gpu_fusion_stage.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02q/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02q/pytest.log
for v in 0 1; do
FFB6D_GATHER_NO_STREAM=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mlp --per-op > gpurun_out/r02q/bench_$v.json 2> gpurun_out/r02q/bench_$v.err
echo "--- NO_STREAM=$v"; grep "gather_max" gpurun_out/r02q/bench_$v.err
python -c "
import json
d=json.loads(open('gpurun_out/r02q/bench_$v.json').read().strip().splitlines()[-1])
print('pass', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['digest_ok'], d['reference_digest_ok'], 'gather', d['compute']['gather_ms_per_step'], 'knn', d['compute']['knn_ms_per_step'])
"
done
