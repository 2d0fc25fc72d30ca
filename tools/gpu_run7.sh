#!/bin/bash
# one B200: two-target walk, CTA size (L1 sharing between adjacent groups)
set -u
O=gpurun_out
mkdir -p $O
: > $O/tree_bench7.log
for v in 2160 4160 8160 4163 8163 8123; do
  echo "== B200_WALK_VARIANT=$v" >> $O/tree_bench7.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread >> $O/tree_bench7.log 2>&1
done
for v in 4160 8160; do
  echo "== B200_WALK_VARIANT=$v (2^24)" >> $O/tree_bench7.log
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread --n 16777216 >> $O/tree_bench7.log 2>&1
done
B200_WALK_VARIANT=8160 ncu --set full --clock-control none --import-source on -k regex:walk_warp -c 1 -f -o $O/r2_walk_v8160 \
      python tools/tree_bench.py --no-thread > $O/ncu_walk_v8160.log 2>&1; echo "ncu 8160 rc=$?" >> $O/run7.log
grep -v "^n=\|lane use" $O/tree_bench7.log
