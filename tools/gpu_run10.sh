#!/bin/bash
# one B200: build graph with the deep levels behind IF nodes
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tree.py tests/test_gpu_forest.py tests/test_gpu_tree_fixed.py tests/test_gpu_host_plugin.py tests/test_gpu_leapfrog.py -m gpu -x -q > $O/pytest_gpu10.log 2>&1; echo "pytest rc=$?" > $O/run10.log
: > $O/tree_bench10.log
for e in 0 1; do
  for d in uniform box clustered; do
    if [ $e = 1 ]; then export B200_NO_COND=1; else unset B200_NO_COND; fi
    echo "== NO_COND=$e $d" >> $O/tree_bench10.log
    python tools/tree_bench.py --no-thread --dist $d >> $O/tree_bench10.log 2>&1
  done
done
unset B200_NO_COND
python tools/tree_bench.py --no-thread --n 16384 >> $O/tree_bench10.log 2>&1
python tools/tree_bench.py --no-thread --n 16777216 >> $O/tree_bench10.log 2>&1
python tools/part_build_time.py --parts 8 >> $O/tree_bench10.log 2>&1
cat $O/run10.log; tail -4 $O/pytest_gpu10.log; grep "^==\|build" $O/tree_bench10.log
