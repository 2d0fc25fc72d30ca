#!/bin/bash
set -u
O=gpurun_out
: > $O/tree_bench12.log
for v in 0 5 6 4; do
  echo "== B200_WALK_VARIANT=$v" >> $O/tree_bench12.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread >> $O/tree_bench12.log 2>&1
done
for v in 0 6; do
  echo "== B200_WALK_VARIANT=$v 2^24" >> $O/tree_bench12.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread --n 16777216 >> $O/tree_bench12.log 2>&1
done
B200_WALK_VARIANT=6 timeout 900 python -m pytest tests/test_gpu_tree.py tests/test_gpu_forest.py -m gpu -x -q > $O/pytest_gpu12.log 2>&1; echo "pytest(6) rc=$?" >> $O/tree_bench12.log
grep "^==\|walk\[\|rc=" $O/tree_bench12.log | cut -c1-75; tail -3 $O/pytest_gpu12.log
