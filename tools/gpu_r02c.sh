#!/bin/bash
# round-2 GPU session C: K=1 tile kernel + everything so far: parity, bench, ncu of the new kernels
mkdir -p gpurun_out/r02c
python -m pytest tests/test_gpu_knn.py tests/test_gpu_pass.py -m gpu -q -x --timeout 1200 -p no:cacheprovider > gpurun_out/r02c/pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02c/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-mlp > gpurun_out/r02c/bench.json 2> gpurun_out/r02c/bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02c/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','digest_ok','reference_digest_ok','gpu_launches')})
print('e2e',d['e2e']['value'],'pass',d['pass_roofline']['frac'])
for k,v in d['compute']['families'].items(): print(k, round(v['ms_per_step'],3))
print(d['compute']['knn_ms_per_step'], d['compute']['gather_ms_per_step'])
"
python tools/kernel_times.py 32 3 > gpurun_out/r02c/kernel_times.txt 2>&1; head -24 gpurun_out/r02c/kernel_times.txt
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'grid_build_kernel|grid_search_group_kernel|grid_search_k1_tile_kernel|grid_search_kernel|gather' -s 50 -c 50 \
  -o gpurun_out/r02c/pass python tools/ncu_pass.py 2 > gpurun_out/r02c/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02c/ncu.log
