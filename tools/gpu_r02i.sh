#!/bin/bash
# session I: full suite, full bench line, L2-fetch-granularity experiment, training-step kernel breakdown
mkdir -p gpurun_out/r02i
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02i/pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r02i/pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02i/bench.json 2> gpurun_out/r02i/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02i/bench.err
python -c "
import json
d=json.load(open('gpurun_out/r02i/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','digest_ok','reference_digest_ok','gpu_launches')})
print('e2e',d['e2e']['value'],'pass',d['pass_roofline']['frac'])
print('stack', d['fusion_stack']['ms_per_step'], d['fusion_stack']['reference_order_ms_per_step'])
print('lfa', d['lfa_blocks'])
print('mlps', d['fusion_mlps']['ms_per_step'])
"
for g in 32 128; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mlp --l2-fetch $g 2>gpurun_out/r02i/l2_$g.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['roofline']['families']
print('l2fetch $g', round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in f.items()})"; tail -1 gpurun_out/r02i/l2_$g.err
done
python tools/train_bench.py --config 3 --steps 3 --warmup 2 --profile > /dev/null 2> gpurun_out/r02i/train_profile.txt; head -30 gpurun_out/r02i/train_profile.txt
