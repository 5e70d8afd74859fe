#!/bin/bash
# round-2 (session 2) experiment call 2: scheduling variants (fixed tool), lean K=1 tile kernel A/B, gather switches
O=gpurun_out/r02ac; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_knn.py tests/test_gpu_pass.py tests/test_gpu_grid.py -q -x -p no:cacheprovider > $O/pytest_knn.log 2>&1; tail -3 $O/pytest_knn.log
FFB6D_CHOOSE_FIRST=1 FFB6D_LAZY_BUILDS=1 timeout 300 python -m pytest tests/test_gpu_pass.py -q -x -p no:cacheprovider > $O/pytest_pass_sw.log 2>&1; tail -2 $O/pytest_pass_sw.log
timeout 400 python tools/pass_ab.py 32 5 20 > $O/ab_sched.log 2>&1; tail -9 $O/ab_sched.log
FFB6D_K1_TILE_OLD=1 timeout 200 python tools/pass_ab.py 32 5 20 base,both > $O/ab_tile_old.log 2>&1; tail -2 $O/ab_tile_old.log
FFB6D_GATHER_NOALLOC=1 timeout 200 python tools/pass_ab.py 32 5 20 base,both > $O/ab_noalloc.log 2>&1; tail -2 $O/ab_noalloc.log
FFB6D_GATHER_SMEM_PAD=47000 timeout 200 python tools/pass_ab.py 32 5 20 base,both > $O/ab_pad.log 2>&1; tail -2 $O/ab_pad.log
FFB6D_CHOOSE_FIRST=1 FFB6D_LAZY_BUILDS=1 timeout 300 python tools/pass_timeline.py $O/timeline_both.json 32 > $O/timeline.log 2>&1; tail -1 $O/timeline.log
