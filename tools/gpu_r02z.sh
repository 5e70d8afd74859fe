#!/bin/bash
# round-2 validation session: full GPU tests, smoke, the default bench line, ncu launch list + full captures
O=gpurun_out/r02z; mkdir -p $O; R=/tmp/ncu_r02z; mkdir -p $R
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > $O/pytest.log 2>&1
echo "pytest rc=$?" | tee -a $O/pytest.log; tail -6 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py > $O/bench.json 2> $O/bench.err
echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','digest_ok','reference_digest_ok','gpu_launches')})
print('e2e',d['e2e']['value'],'pass',d['pass_roofline']['frac'], 'roofline', d['roofline']['kernel'], d['roofline']['frac'])
for k in ('fusion_mlps','fusion_stack','lfa_blocks','pose_voting','gpu_torch_reference','host_api','cpu_baseline','clocks'):
    print(k, json.dumps(d.get(k))[:600])
"
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"; cut -c1-400 $O/bench_reference.json
# launch list of the bench command (per-launch durations; cold-cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-mlp > $O/launches.log 2>&1; echo "launches rc=$?"
# full captures
timeout 600 ncu --set full --clock-control none \
  -k regex:'grid_build_kernel|grid_search_k1_tile_kernel|grid_search_group_kernel|grid_search_kernel|gather1_ncs|gather_max_ncs' -c 90 \
  -o $R/pass python tools/ncu_pass.py 1 > $O/ncu_pass.log 2>&1; echo "ncu pass rc=$?"; tail -2 $O/ncu_pass.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fusion_mlp_packed_kernel' -s 2 -c 1 \
  -o $R/mlp_mt2 python tools/ncu_mlp.py > $O/ncu_mlp.log 2>&1; echo "ncu mlp rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fusion_mlp_packed_kernel' -s 2 -c 1 \
  -o $R/mlp_512 python tools/ncu_mlp.py 256 256 256 19200 > $O/ncu_mlp2.log 2>&1; echo "ncu mlp2 rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'lfa_att_pool_fused' -s 18 -c 6 \
  -o $R/lfa python tools/lfa_times.py 32 > $O/ncu_lfa.log 2>&1; echo "ncu lfa rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'mean_shift_kernel' -c 1 \
  -o $R/pose python tools/ncu_pose.py 4096 30 > $O/ncu_pose.log 2>&1; echo "ncu pose rc=$?"; tail -2 $O/ncu_pose.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fusion_wgrad_kernel|wgrad_small_kernel|bn_stats_kernel|bn_bwd_reduce_kernel|gather1_bwd_runs_kernel|att_pool_bwd_kernel' -c 24 \
  -o $R/train python tools/train_bench.py --config 3 --steps 1 --warmup 0 > $O/ncu_train.log 2>&1; echo "ncu train rc=$?"
for n in pass mlp_mt2 mlp_512 lfa pose train; do [ -f $R/$n.ncu-rep ] && python tools/ncu_table.py $R/$n.ncu-rep > $O/ncu_$n.md; done
cp $R/pose.ncu-rep $R/mlp_mt2.ncu-rep $O/ 2>/dev/null
ls -la $O $R
