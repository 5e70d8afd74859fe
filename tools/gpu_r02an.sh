#!/bin/bash
# session 2: timeline of the captured pass on the committed defaults
O=gpurun_out/r02an; mkdir -p $O
timeout 80 python tools/pass_timeline.py $O/timeline_final.json 32 > $O/timeline.log 2>&1; tail -1 $O/timeline.log
