#!/bin/bash
# session 2, final sanity run of the committed state: full GPU suite + smoke
O=gpurun_out/r02am; mkdir -p $O
python -m pytest tests -m gpu -q --timeout 1200 -p no:cacheprovider > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
