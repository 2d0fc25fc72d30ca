#!/bin/bash
# one B200: persistent two-target walk with the SM-local scheduler
set -u
O=gpurun_out
B200_WALK_VARIANT=5 timeout 1500 python -m pytest tests/test_gpu_tree.py tests/test_gpu_forest.py -m gpu -x -q > $O/pytest_gpu11.log 2>&1; echo "pytest(5) rc=$?" > $O/run11.log
: > $O/tree_bench11.log
for v in 0 5 6 7; do
  echo "== B200_WALK_VARIANT=$v" >> $O/tree_bench11.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread >> $O/tree_bench11.log 2>&1
done
for o in 10 15 16; do
  echo "== B200_WALK_VARIANT=5 OWN=$o" >> $O/tree_bench11.log
  B200_WALK_VARIANT=5 B200_WALK_OWN=$o python tools/tree_bench.py --no-thread >> $O/tree_bench11.log 2>&1
done
for v in 0 5 6; do
  echo "== B200_WALK_VARIANT=$v clustered" >> $O/tree_bench11.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread --dist clustered >> $O/tree_bench11.log 2>&1
  echo "== B200_WALK_VARIANT=$v 2^24" >> $O/tree_bench11.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread --n 16777216 >> $O/tree_bench11.log 2>&1
done
B200_WALK_VARIANT=5 ncu --set full --clock-control none --import-source on -k regex:walk_warp -c 1 -f -o $O/r2_walk_v5 \
      python tools/tree_bench.py --no-thread > $O/ncu_walk_v5.log 2>&1; echo "ncu 5 rc=$?" >> $O/run11.log
cat $O/run11.log; tail -3 $O/pytest_gpu11.log; grep "^==\|walk\[" $O/tree_bench11.log | cut -c1-75
