#!/bin/bash
# round-2 GPU session E: training-mode layers + full suite + K=1 grid scale sweep
mkdir -p gpurun_out/r02e
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/r02e/pytest.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/r02e/pytest.log
for sc in 1.0 1.5 2.0 2.5; do
FFB6D_GRID_SCALE_K1=$sc python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mlp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['compute']['families']
print('k1 scale $sc', round(d['ms_per_step'],3), 'k1', round(f['grid_search_kernel<K=1> (+knn_brute for S<512)']['ms_per_step'],3), d['digest_ok'])"
done
for sc in 0.8 1.25 1.5; do
FFB6D_GRID_SCALE=$sc python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-mlp 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); f=d['compute']['families']
print('k16 scale $sc', round(d['ms_per_step'],3), 'self', round(f['grid_search_warp_kernel<SELF>']['ms_per_step'],3), 'nonself', round(f['grid_search_warp_kernel<non-self>']['ms_per_step'],3), d['digest_ok'])"
done
