#!/bin/bash
set -u
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tree.py tests/test_gpu_forest.py tests/test_gpu_host_plugin.py -m gpu -x -q > $O/pytest_gpu15.log 2>&1; echo "pytest rc=$?" > $O/run15.log
python tools/part_build_time.py --parts 8 > $O/part_build15.log 2>&1
python tools/part_build_time.py --parts 2 >> $O/part_build15.log 2>&1
python tools/forest_walk_time.py --parts 8 --cold --drift 0.15 2>&1 | grep "per-rank" >> $O/part_build15.log
cat $O/run15.log; tail -3 $O/pytest_gpu15.log; cat $O/part_build15.log
