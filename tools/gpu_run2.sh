#!/bin/bash
# round-2 GPU session 2 (two B200): full -m gpu suite incl. forest tests, bench at 1 and 2 GPUs
set -u
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu2.log 2>&1; echo "pytest rc=$?" > $O/run2.log
python bench.py --no-cpu --no-c5 --steps 3 > $O/bench_r2b_1gpu.json 2> $O/bench_r2b_1gpu.err; echo "bench1 rc=$?" >> $O/run2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_r2b_2gpu.json 2> $O/bench_r2b_2gpu.err; echo "bench2 rc=$?" >> $O/run2.log
cat $O/run2.log; tail -5 $O/pytest_gpu2.log; tail -3 $O/bench_r2b_2gpu.err
