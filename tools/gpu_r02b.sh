#!/bin/bash
# round-2 GPU session B: new grid build (cluster kernel) + half-warp K<=16 search: parity, bench, variants
mkdir -p gpurun_out/r02b
python -m pytest tests/test_gpu_knn.py tests/test_gpu_pass.py -m gpu -q -x --timeout 1200 -p no:cacheprovider > gpurun_out/r02b/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02b/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-mlp > gpurun_out/r02b/bench.json 2> gpurun_out/r02b/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02b/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','digest_ok','reference_digest_ok','gpu_launches')})
print('e2e',d['e2e']['value'],'pass',d['pass_roofline']['frac'])
for k,v in d['compute']['families'].items(): print(k, round(v['ms_per_step'],3))
print(d['compute']['knn_ms_per_step'], d['compute']['gather_ms_per_step'])
"
for v in "4 4" "3 3" "2 4"; do set -- $v
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mlp --streams $1 --gather-streams $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('streams $1 $2', d['ms_per_step'], d['e2e']['ms_per_step'])"
done
python tools/kernel_times.py 32 3 > gpurun_out/r02b/kernel_times.txt 2>&1; head -30 gpurun_out/r02b/kernel_times.txt
