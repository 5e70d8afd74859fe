#!/bin/bash
# round-2 GPU session D: staged K=1 tile kernel, group kernel at 5 CTAs/SM, f-3 fused stage
mkdir -p gpurun_out/r02d
python -m pytest tests/test_gpu_knn.py tests/test_gpu_pass.py tests/test_gpu_fusion_stage.py tests/test_gpu_fusion_mlp.py tests/test_gpu_lfa.py -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02d/pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r02d/pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02d/bench.json 2> gpurun_out/r02d/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02d/bench.err
python -c "
import json
d=json.load(open('gpurun_out/r02d/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','digest_ok','reference_digest_ok','gpu_launches')})
print('e2e',d['e2e']['value'],'pass',d['pass_roofline']['frac'])
for k,v in d['compute']['families'].items(): print(k, round(v['ms_per_step'],3))
print(d['compute']['knn_ms_per_step'], d['compute']['gather_ms_per_step'])
print('mlps', d['fusion_mlps']['ms_per_step'], d['fusion_mlps']['fp32_equiv_tflops'])
print('stack', d['fusion_stack'])
"
python tools/kernel_times.py 32 3 > gpurun_out/r02d/kernel_times.txt 2>&1; head -22 gpurun_out/r02d/kernel_times.txt
