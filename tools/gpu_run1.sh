#!/bin/bash
# round-2 GPU session 1 (one B200): goldens, tests, bench, launch list, ncu captures
set -u
O=gpurun_out
mkdir -p $O
python tests/golden/make_golden_gpu.py > $O/golden.log 2>&1 && cp $O/ref_gpu_energy.npz tests/golden/
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/run1.log
python bench.py > $O/bench_r2_1gpu.json 2> $O/bench_r2_1gpu.err; echo "bench rc=$?" >> $O/run1.log
python bench.py --workload tree --no-cpu > $O/bench_r2_tree_1gpu.json 2> $O/bench_r2_tree_1gpu.err; echo "bench tree rc=$?" >> $O/run1.log
python tools/leapfrog_time.py > $O/leapfrog_tile.json 2>> $O/run1.log
B200_LEAPFROG=v1 python tools/leapfrog_time.py > $O/leapfrog_v1.json 2>> $O/run1.log
python tools/tree_bench.py > $O/tree_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-c5 > $O/ncu_bench.log 2>&1; echo "ncu list rc=$?" >> $O/run1.log
ncu --set full --clock-control none --import-source on -k regex:walk_warp_kernel -c 1 -f -o $O/r2_walk \
    python tools/tree_bench.py > $O/ncu_walk.log 2>&1; echo "ncu walk rc=$?" >> $O/run1.log
ncu --set full --clock-control none --import-source on -k regex:leapfrog_kernel -c 1 -f -o $O/r2_leapfrog \
    python tools/leapfrog_time.py > $O/ncu_leapfrog.log 2>&1; echo "ncu leapfrog rc=$?" >> $O/run1.log
B200_LEAPFROG=v1 ncu --set full --clock-control none -k regex:leapfrog_v1_kernel -c 1 -f -o $O/r2_leapfrog_v1 \
    python tools/leapfrog_time.py > $O/ncu_leapfrog_v1.log 2>&1; echo "ncu leapfrog v1 rc=$?" >> $O/run1.log
cat $O/run1.log; tail -3 $O/pytest_gpu.log
