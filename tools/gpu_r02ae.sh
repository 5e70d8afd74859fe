#!/bin/bash
# session 2, call 4: far-query sentinels in the lean K=1 tile kernel -- full GPU suite, A/B vs the old tile kernel, warps per CTA
O=gpurun_out/r02ae; mkdir -p $O
python -m pytest tests -m gpu -q -x --timeout 1200 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
for w in 8 2 1; do
FFB6D_K1_TILE_WARPS=$w timeout 200 python tools/pass_ab.py 32 5 20 base,choose_first > $O/ab_tilewarps_$w.log 2>&1; echo "tile warps $w"; tail -2 $O/ab_tilewarps_$w.log
done
FFB6D_K1_TILE_OLD=1 timeout 200 python tools/pass_ab.py 32 5 20 base,choose_first > $O/ab_tile_old.log 2>&1; echo "old tile kernel"; tail -2 $O/ab_tile_old.log
timeout 300 python tools/pass_timeline.py $O/timeline_base.json 32 > $O/timeline.log 2>&1; tail -1 $O/timeline.log
