#!/bin/bash
# one B200: walk variants with the successor requested right after the vote (PF=3)
set -u
O=gpurun_out
mkdir -p $O
B200_WALK_VARIANT=2143 timeout 1500 python -m pytest tests/test_gpu_tree.py tests/test_gpu_forest.py -m gpu -x -q > $O/pytest_gpu6.log 2>&1; echo "pytest(2143) rc=$?" > $O/run6.log
: > $O/tree_bench6.log
for v in 90 93 103 83 2140 2160 2123 2143 2163; do
  echo "== B200_WALK_VARIANT=$v" >> $O/tree_bench6.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread >> $O/tree_bench6.log 2>&1
done
for v in 93 2143 2163; do
  echo "== B200_WALK_VARIANT=$v (2^24)" >> $O/tree_bench6.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread --n 16777216 >> $O/tree_bench6.log 2>&1
done
B200_WALK_VARIANT=2143 ncu --set full --clock-control none --import-source on -k regex:walk_warp -c 1 -f -o $O/r2_walk_v2143 \
      python tools/tree_bench.py --no-thread > $O/ncu_walk_v2143.log 2>&1; echo "ncu 2143 rc=$?" >> $O/run6.log
cat $O/run6.log; tail -3 $O/pytest_gpu6.log; grep -v "^n=\|lane use" $O/tree_bench6.log
