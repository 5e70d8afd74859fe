#!/bin/bash
# session 2 validation A: full GPU suite, smoke, ncu launch list + full captures of one sequential pass
O=gpurun_out/r02aj; mkdir -p $O; R=/tmp/ncu_r02aj; mkdir -p $R
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 ncu --set full --clock-control none \
  -k regex:'grid_build_kernel|grid_search_k1_tile_kernel|grid_far_fill_kernel|grid_search_group_kernel|grid_search_kernel|gather1_ncs|gather_max_ncs' -c 90 \
  -o $R/pass python tools/ncu_pass.py 1 > $O/ncu_pass.log 2>&1; echo "ncu pass rc=$?"; tail -2 $O/ncu_pass.log
python tools/ncu_table.py $R/pass.ncu-rep > $O/ncu_pass.md; wc -l $O/ncu_pass.md
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-mlp > $O/launches.log 2>&1; echo "launches rc=$?"
ls -la $O
