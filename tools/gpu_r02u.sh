#!/bin/bash
O=gpurun_out/r02u; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_pass.py tests/test_gpu_schedule.py tests/test_gpu_backproject.py tests/test_gpu_knn.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $O/pytest.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-mlp --per-op > $O/bench.json 2> $O/bench.err
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print('pass', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['digest_ok'], d['reference_digest_ok'], 'gather', d['compute']['gather_ms_per_step'], 'knn', d['compute']['knn_ms_per_step'], d['gpu_launches'])
"
grep "knn" $O/bench.err | head -30
