#!/bin/bash
# 8-GPU validation of the sharded paths (run under: gpurun --gpus 8 -- bash tools/scale8.sh)
O=gpurun_out
mkdir -p $O
run() { # name, nproc, args...
  name=$1; np=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
     bench.py --gpus $np "$@" > $O/$name.json 2> $O/$name.err
  echo "$name rc=$?" >> $O/scale8.log
}
: > $O/scale8.log
run r2_bench_8gpu 8 --steps 5 --warmup 3
run r2_bench_4gpu 4 --steps 5 --warmup 3 --no-c5
run r2_bench_2gpu 2 --steps 5 --warmup 3 --no-c5
tests/host/_bin/shard_test > $O/r2_shard_test_8gpu_box.log 2>&1; echo "shard_test rc=$?" >> $O/scale8.log
timeout 600 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > $O/pytest_peer8.log 2>&1; echo "pytest peer rc=$?" >> $O/scale8.log
cat $O/scale8.log; tail -3 $O/r2_shard_test_8gpu_box.log; tail -3 $O/pytest_peer8.log
