#!/bin/bash
# session 2, call 7: unrolled grid build (A/B vs previous numbers), source-level captures of the other heavy kernels
O=gpurun_out/r02ah; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_knn.py tests/test_gpu_grid.py tests/test_gpu_pass.py -q -x -p no:cacheprovider > $O/pytest.log 2>&1; tail -2 $O/pytest.log
timeout 300 python tools/pass_ab.py 32 5 20 base,self_first > $O/ab.log 2>&1; grep median $O/ab.log
timeout 600 ncu --set full --clock-control none --import-source on \
  -k regex:'gather_max_ncs_staged_kernel|gather_max_ncs_klane_kernel|gather1_ncs_staged_v4_kernel|gather1_ncs_direct_kernel|grid_search_kernel|grid_build_kernel' \
  -o $O/heavy python tools/ncu_pass.py 1 > $O/ncu_heavy.log 2>&1; echo "ncu rc=$?"
ls -la $O
