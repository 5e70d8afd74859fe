#!/bin/bash
# session 2, call 5: head-of-pass scheduling variants (self search first, build stream, async set copies)
O=gpurun_out/r02af; mkdir -p $O
timeout 500 python tools/pass_ab.py 32 5 20 > $O/ab_sched.log 2>&1; tail -10 $O/ab_sched.log
FFB6D_SELF_FIRST=1 FFB6D_ASYNC_SETS=1 FFB6D_BUILD_STREAMS=1 timeout 300 python -m pytest tests/test_gpu_pass.py tests/test_gpu_knn.py -q -x -p no:cacheprovider > $O/pytest_sw.log 2>&1; tail -2 $O/pytest_sw.log
FFB6D_SELF_FIRST=1 FFB6D_ASYNC_SETS=1 FFB6D_BUILD_STREAMS=1 timeout 300 python tools/pass_timeline.py $O/timeline_bs1_self_async.json 32 > $O/timeline.log 2>&1; tail -1 $O/timeline.log
