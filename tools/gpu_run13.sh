#!/bin/bash
set -u
O=gpurun_out
python tools/forest_walk_time.py --parts 4 --cold --drift 0.02 --single-only > $O/forest_walk_time6.log 2>&1
python tools/forest_walk_time.py --parts 8 --cold --drift 0.15 >> $O/forest_walk_time6.log 2>&1
python tools/tree_bench.py --no-thread >> $O/forest_walk_time6.log 2>&1
grep -v "list in" $O/forest_walk_time6.log
