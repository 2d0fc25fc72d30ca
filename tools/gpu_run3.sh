#!/bin/bash
# one B200: tree/forest/host tests after the compact walk records, tree timings
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_tree.py tests/test_gpu_forest.py tests/test_gpu_tree_fixed.py tests/test_gpu_host_plugin.py tests/test_gpu_leapfrog.py tests/test_gpu_energy.py -m gpu -x -q > $O/pytest_gpu3.log 2>&1; echo "pytest rc=$?" > $O/run3.log
python tools/tree_bench.py > $O/tree_bench3.log 2>&1
python tools/tree_bench.py --dist box >> $O/tree_bench3.log 2>&1
python tools/tree_bench.py --n 16777216 >> $O/tree_bench3.log 2>&1
python bench.py --no-cpu --no-c5 --steps 3 > $O/bench_r2c_1gpu.json 2> $O/bench_r2c_1gpu.err; echo "bench1 rc=$?" >> $O/run3.log
cat $O/run3.log; tail -3 $O/pytest_gpu3.log; cat $O/tree_bench3.log
