#!/bin/bash
# session 2, call 3: full GPU suite after the schedule changes + ncu source-level captures of the K=1 tile kernels
O=gpurun_out/r02ad; mkdir -p $O; R=$O
python -m pytest tests -m gpu -q -x --timeout 1200 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'grid_search_k1_tile2_kernel' -c 2 \
  -o $R/tile2 python tools/ncu_pass.py 1 > $O/ncu_tile2.log 2>&1; echo "ncu tile2 rc=$?"
FFB6D_K1_TILE_OLD=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'grid_search_k1_tile_kernel' -c 2 \
  -o $R/tile1 python tools/ncu_pass.py 1 > $O/ncu_tile1.log 2>&1; echo "ncu tile1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'grid_search_group_kernel' -c 1 \
  -o $R/group python tools/ncu_pass.py 1 > $O/ncu_group.log 2>&1; echo "ncu group rc=$?"
ls -la $O
for w in 8 2 1; do
FFB6D_K1_TILE_WARPS=$w timeout 200 python tools/pass_ab.py 32 5 20 base,choose_first > $O/ab_tilewarps_$w.log 2>&1; echo "tile warps $w"; tail -2 $O/ab_tilewarps_$w.log
done
