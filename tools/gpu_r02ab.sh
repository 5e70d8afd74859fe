#!/bin/bash
# round-2 (session 2) experiment call 1: timeline of the captured pass + A/B of scheduling variants and two
# gather launch switches (L1 no-allocate loads, occupancy cap)
O=gpurun_out/r02ab; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 300 python tools/pass_timeline.py $O/timeline_base.json 32 > $O/timeline.log 2>&1; tail -1 $O/timeline.log
timeout 400 python tools/pass_ab.py 32 5 20 > $O/ab_sched.log 2>&1; cat $O/ab_sched.log | tail -20
FFB6D_GATHER_NOALLOC=1 timeout 200 python tools/pass_ab.py 32 5 20 base,both > $O/ab_noalloc.log 2>&1; tail -4 $O/ab_noalloc.log
FFB6D_GATHER_SMEM_PAD=47000 timeout 200 python tools/pass_ab.py 32 5 20 base,both > $O/ab_pad.log 2>&1; tail -4 $O/ab_pad.log
FFB6D_GATHER_NOALLOC=1 FFB6D_GATHER_SMEM_PAD=47000 timeout 200 python tools/pass_ab.py 32 5 20 base,both > $O/ab_noalloc_pad.log 2>&1; tail -4 $O/ab_noalloc_pad.log
FFB6D_GATHER_NOALLOC=1 FFB6D_CHOOSE_FIRST=1 FFB6D_LAZY_BUILDS=1 timeout 300 python -m pytest tests/test_gpu_pass.py tests/test_gpu_gather.py -q -x -p no:cacheprovider > $O/pytest_switches.log 2>&1; tail -3 $O/pytest_switches.log
