#!/bin/bash
O=gpurun_out/r02aa; mkdir -p $O
for v in ready size ready size; do
FFB6D_GATHER_ORDER=$v python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-mlp > $O/bench_$v.json 2> $O/bench_$v.err
python -c "
import json
d=json.loads(open('$O/bench_$v.json').read().strip().splitlines()[-1])
print('ORDER=$v pass', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['digest_ok'], d['reference_digest_ok'])
"
done
