#!/bin/bash
# 2 GPUs: DDP training bench (configs 3, 4), bench.py N=2, model test
mkdir -p gpurun_out/r02h
python -m pytest tests/test_gpu_model.py tests/test_gpu_pass.py::test_two_devices_in_one_process -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -5
for c in 3 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_bench.py --config $c --steps 5 --warmup 2 > gpurun_out/r02h/train${c}_n2.json 2> gpurun_out/r02h/train${c}_n2.err
echo "train$c n2 rc=$?"; cat gpurun_out/r02h/train${c}_n2.json; tail -3 gpurun_out/r02h/train${c}_n2.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-mlp > gpurun_out/r02h/bench_n2.json 2> gpurun_out/r02h/bench_n2.err
echo "bench n2 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02h/bench_n2.json')); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['digest_ok'])"
