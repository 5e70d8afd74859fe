#!/bin/bash
# round-2 GPU session A: tests, bench, ncu captures of the round-1 kernels the verdict named
mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a/smi.txt
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02a/pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02a/pytest.log
tail -15 gpurun_out/r02a/pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02a/bench.json 2> gpurun_out/r02a/bench.err
echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02a/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','digest_ok','reference_digest_ok','gpu_launches')})
print('e2e',d['e2e']['value'],'pass',d['pass_roofline']['frac'])
print(json.dumps(d['compute'],indent=0)[:1500])
print(json.dumps(d['roofline']['families'],indent=0)[:1500])
print(d['gpu_torch_reference']); print(d['host_api']); print(d['cpu_baseline'])
"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'grid_prepare_kernel|gather1_ncs_direct_kernel|grid_search_warp_kernel|grid_search_kernel' -s 25 -c 25 \
  -o gpurun_out/r02a/r01kernels python tools/ncu_pass.py 2 > gpurun_out/r02a/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02a/ncu.log
