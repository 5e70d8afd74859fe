#!/bin/bash
# N GPUs (first argument): DDP training bench (configs 3, 4) and bench.py under torchrun
N=${1:-8}; O=gpurun_out/r02t; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for c in 3 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$c tools/train_bench.py --config $c --steps 10 --warmup 3 > $O/train_config${c}_n$N.json 2> $O/train_config${c}_n$N.err
echo "train$c n$N rc=$?"; grep '^{' $O/train_config${c}_n$N.json | cut -c1-600; tail -2 $O/train_config${c}_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-mlp > $O/bench_n$N.json 2> $O/bench_n$N.err
echo "bench n$N rc=$?"; python -c "
import json; d=json.loads([l for l in open('$O/bench_n$N.json') if l.startswith('{')][-1]); print(d['value'], d['ms_per_step'], d['n_gpus'], d['e2e']['value'], d['digest_ok'], d['clocks'])"
