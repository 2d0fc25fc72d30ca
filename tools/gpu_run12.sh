#!/bin/bash
set -u
O=gpurun_out
: > $O/tree_bench12.log
for v in 0 5 6 7 8 4; do
  echo "== B200_WALK_VARIANT=$v" >> $O/tree_bench12.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread >> $O/tree_bench12.log 2>&1
done
for v in 0 5; do
  echo "== B200_WALK_VARIANT=$v 2^24" >> $O/tree_bench12.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread --n 16777216 >> $O/tree_bench12.log 2>&1
done
bash tools/bounds_check.sh > $O/bounds_check.log 2>&1; echo "bounds_check rc=$?" >> $O/tree_bench12.log
grep "^==\|walk\[\|rc=" $O/tree_bench12.log | cut -c1-75; tail -3 $O/bounds_check.log
