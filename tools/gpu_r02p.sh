#!/bin/bash
mkdir -p gpurun_out/r02p
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mlp --per-op > gpurun_out/r02p/bench.json 2> gpurun_out/r02p/bench.err
grep "ms/step" gpurun_out/r02p/bench.err | head -60
python -c "
import json
d=json.loads(open('gpurun_out/r02p/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['digest_ok'], d['reference_digest_ok'])
print(json.dumps(d['compute'])[:1500])
"
