#!/bin/bash
# session 2, last call: the default bench line of the committed defaults + in-process A/B of the cld_interp derivation
O=gpurun_out/r02al; mkdir -p $O
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','digest_ok','reference_digest_ok','gpu_launches')})
print('e2e',d['e2e']['value'],'pass',d['pass_roofline']['frac'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'])
print('knn', d['compute']['knn_ms_per_step'], 'gather', d['compute']['gather_ms_per_step'])
"
timeout 200 python tools/pass_ab.py 32 5 20 > $O/ab.log 2>&1; grep "median\|MISMATCH" $O/ab.log
