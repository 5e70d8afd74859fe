#!/bin/bash
mkdir -p gpurun_out/r02l
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lfa_att_pool_fused' -s 18 -c 6 \
  -o gpurun_out/r02l/lfa python tools/lfa_times.py 32 > gpurun_out/r02l/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02l/ncu.log
