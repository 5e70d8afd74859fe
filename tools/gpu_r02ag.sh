#!/bin/bash
# session 2, call 6: grid builds on a high-priority stream (their CTAs are placed ahead of a running big grid's pending CTAs)
O=gpurun_out/r02ag; mkdir -p $O
timeout 500 python tools/pass_ab.py 32 5 20 > $O/ab_sched.log 2>&1; grep median $O/ab_sched.log
FFB6D_SELF_FIRST=1 FFB6D_ASYNC_SETS=1 FFB6D_BUILD_STREAMS=1 FFB6D_BUILD_PRIO=-1 timeout 300 python tools/pass_timeline.py $O/timeline_bs1hi_self_async.json 32 > $O/timeline.log 2>&1; tail -1 $O/timeline.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'grid_build_kernel' -c 1 \
  -o $O/build python tools/ncu_pass.py 1 > $O/ncu_build.log 2>&1; echo "ncu build rc=$?"
