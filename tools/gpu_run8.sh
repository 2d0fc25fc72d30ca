#!/bin/bash
# one B200: full GPU suite + bench after the depth-first walk tables / two-target walk
set -u
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu8.log 2>&1; echo "pytest rc=$?" > $O/run8.log
python bench.py --no-c5 --steps 3 > $O/bench_r2d_1gpu.json 2> $O/bench_r2d_1gpu.err; echo "bench rc=$?" >> $O/run8.log
python tools/tree_bench.py --dist clustered > $O/tree_bench8.log 2>&1
python tools/tree_bench.py --dist box >> $O/tree_bench8.log 2>&1
cat $O/run8.log; tail -4 $O/pytest_gpu8.log; cat $O/tree_bench8.log
