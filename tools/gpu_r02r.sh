#!/bin/bash
mkdir -p gpurun_out/r02r
for v in 0 1 0 1; do
FFB6D_GATHER_NO_STREAM=$v python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-mlp > gpurun_out/r02r/bench_$v.json 2> gpurun_out/r02r/bench_$v.err
python -c "
import json
d=json.loads(open('gpurun_out/r02r/bench_$v.json').read().strip().splitlines()[-1])
print('NO_STREAM=$v pass', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['digest_ok'], d['reference_digest_ok'], 'gather', d['compute']['gather_ms_per_step'], 'knn', d['compute']['knn_ms_per_step'])
"
done
