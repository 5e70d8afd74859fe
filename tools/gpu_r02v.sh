#!/bin/bash
# final validation of the round: full GPU tests, smoke, the default bench line, reference arm, stress sweep
O=gpurun_out/r02v; mkdir -p $O
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/pytest.log; tail -6 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','digest_ok','reference_digest_ok','gpu_launches')})
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'pass',d['pass_roofline']['frac'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'])
print(json.dumps(d['compute'])[:1500])
for k in ('fusion_mlps','fusion_stack','lfa_blocks','pose_voting','clocks'):
    print(k, json.dumps(d.get(k))[:400])
"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 900 python tools/sweep.py --steps 10 > $O/sweep.md 2> $O/sweep.err; cat $O/sweep.md; cp gpurun_out/sweep.jsonl $O/
