#!/bin/bash
O=gpurun_out/r02w; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pass.py -m gpu -q --timeout 600 -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $O/pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'grid_search_group_kernel' -c 1 \
  -o $O/group_self python tools/ncu_pass.py 1 > $O/ncu1.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'grid_search_k1_tile_kernel' -s 2 -c 1 \
  -o $O/k1_tile python tools/ncu_pass.py 1 > $O/ncu2.log 2>&1; echo "rc=$?"
ls -la $O
