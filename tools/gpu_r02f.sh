#!/bin/bash
# round-2 GPU session F: model-level parity (eval + train), training bench configs 3/4 on 1 GPU, stream experiments
mkdir -p gpurun_out/r02f
python -m pytest tests/test_gpu_model.py tests/test_gpu_train.py -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02f/pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/r02f/pytest.log
python tools/train_bench.py --config 3 --steps 5 --warmup 2 > gpurun_out/r02f/train3_n1.json 2> gpurun_out/r02f/train3.err; echo "train3 rc=$?"; cat gpurun_out/r02f/train3_n1.json; tail -5 gpurun_out/r02f/train3.err
python tools/train_bench.py --config 4 --steps 5 --warmup 2 > gpurun_out/r02f/train4_n1.json 2> gpurun_out/r02f/train4.err; echo "train4 rc=$?"; cat gpurun_out/r02f/train4_n1.json; tail -5 gpurun_out/r02f/train4.err
for pr in "0,0" "-1,0" "0,-1"; do
FFB6D_STREAM_PRIO=$pr python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mlp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('prio $pr', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['digest_ok'])"
done
for v in "1 1" "2 3" "3 2" "2 6"; do set -- $v
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mlp --streams $1 --gather-streams $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('streams $1 $2', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))"
done
