#!/bin/bash
# session 2, call 8: staged gathers -- prefetched index loads, threads per CTA
O=gpurun_out/r02ai; mkdir -p $O
FFB6D_GATHER_THREADS=512 FFB6D_GATHER_THREADS_K1=512 timeout 300 python -m pytest tests/test_gpu_gather.py tests/test_gpu_pass.py -q -x -p no:cacheprovider > $O/pytest.log 2>&1; tail -2 $O/pytest.log
for cfg in "256 256" "512 256" "1024 256" "512 512" "256 512"; do set -- $cfg
FFB6D_GATHER_THREADS=$1 FFB6D_GATHER_THREADS_K1=$2 timeout 200 python tools/pass_ab.py 32 5 20 base > $O/ab_$1_$2.log 2>&1; echo "threads K16=$1 K1=$2: $(grep median $O/ab_$1_$2.log)"
done
