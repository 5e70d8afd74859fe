#!/bin/bash
# session 2 validation B (final code): full GPU suite, the default bench line, and the same pass without the
# cld_interp derivation (A/B)
O=gpurun_out/r02ak; mkdir -p $O
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
SECONDS=0
python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$? in ${SECONDS}s"
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','digest_ok','reference_digest_ok','gpu_launches')})
print('e2e',d['e2e']['value'],'pass',d['pass_roofline']['frac'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'])
print('knn', d['compute']['knn_ms_per_step'], 'gather', d['compute']['gather_ms_per_step'])
"
SECONDS=0
FFB6D_SUBSET_NN=0 python bench.py --no-cpu-baseline --no-mlp > $O/bench_nosubset.json 2> $O/bench_nosubset.err; echo "bench(no subset) rc=$? in ${SECONDS}s"
python -c "
import json
d=json.loads(open('$O/bench_nosubset.json').read().strip().splitlines()[-1])
print('no-subset', {k:d[k] for k in ('value','ms_per_step','digest_ok','reference_digest_ok','gpu_launches')}, 'knn', d['compute']['knn_ms_per_step'])
"
