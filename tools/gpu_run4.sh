#!/bin/bash
# one B200: parity of the depth-first walk tables with the one- and the two-target walk, then the walk variants
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_tree.py tests/test_gpu_forest.py tests/test_gpu_tree_fixed.py tests/test_gpu_host_plugin.py -m gpu -x -q > $O/pytest_gpu4.log 2>&1; echo "pytest(default) rc=$?" > $O/run4.log
B200_WALK_VARIANT=2141 timeout 1500 python -m pytest tests/test_gpu_tree.py tests/test_gpu_forest.py tests/test_gpu_host_plugin.py -m gpu -x -q > $O/pytest_gpu4b.log 2>&1; echo "pytest(2141) rc=$?" >> $O/run4.log
: > $O/tree_bench4.log
for v in 90 91 92 100 101 81 2120 2121 2140 2141 2160 2161; do
  echo "== B200_WALK_VARIANT=$v" >> $O/tree_bench4.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread >> $O/tree_bench4.log 2>&1
done
for v in 91 2121 2141; do
  echo "== B200_WALK_VARIANT=$v (2^24)" >> $O/tree_bench4.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread --n 16777216 >> $O/tree_bench4.log 2>&1
done
B200_WALK_VARIANT=2141 python tools/tree_bench.py --dist box >> $O/tree_bench4.log 2>&1
B200_WALK_VARIANT=2141 python tools/tree_bench.py --dist clustered >> $O/tree_bench4.log 2>&1
cat $O/run4.log; tail -3 $O/pytest_gpu4.log; tail -3 $O/pytest_gpu4b.log; grep -v "^n=" $O/tree_bench4.log
