#!/bin/bash
O=gpurun_out/r02s; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_pose.py -m gpu -q --timeout 300 -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 300 python tools/pose_times.py 2>&1 | tee $O/pose_times.txt
for c in 3 4; do
python tools/train_bench.py --config $c --steps 10 --warmup 3 > $O/train_config${c}_n1.json 2> $O/train_config${c}_n1.err; cut -c1-700 $O/train_config${c}_n1.json
done
python tools/train_bench.py --config 3 --steps 3 --warmup 2 --profile 2> $O/train_profile.txt > /dev/null; sed -n 3,14p $O/train_profile.txt
