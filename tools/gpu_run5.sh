#!/bin/bash
# one B200: ncu of the two-target walk (2160) and of the one-target walk on depth-first tables (90)
set -u
O=gpurun_out
mkdir -p $O
for v in 90 2160; do
  B200_WALK_VARIANT=$v python tools/tree_bench.py --no-thread > $O/tree_bench5_$v.log 2>&1
  B200_WALK_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:walk_warp -c 1 -f -o $O/r2_walk_v$v \
      python tools/tree_bench.py --no-thread > $O/ncu_walk_v$v.log 2>&1; echo "ncu $v rc=$?" >> $O/run5.log
done
cat $O/run5.log $O/tree_bench5_*.log
