#!/bin/bash
# 8-GPU validation of the sharded paths (run under: gpurun --gpus 8 -- bash tools/scale8.sh)
run() { # name, nproc, args...
  name=$1; np=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
     bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?" >> gpurun_out/scale8.log
}
: > gpurun_out/scale8.log
run tree_16m_kdk_8gpu 8 --steps 5 --warmup 3 --workload tree --particles 16777216 --kdk --no-cpu
run tree_16m_kdk_4gpu 4 --steps 3 --warmup 3 --workload tree --particles 16777216 --kdk --no-cpu
run tree_16m_kdk_2gpu 2 --steps 3 --warmup 3 --workload tree --particles 16777216 --kdk --no-cpu
run direct_1m_8gpu 8 --steps 5 --warmup 3 --no-cpu
run tree_1m_8gpu 8 --steps 5 --warmup 3 --workload tree --no-cpu
tests/host/_bin/shard_test > gpurun_out/shard_test8.log 2>&1; echo "shard_test rc=$?" >> gpurun_out/scale8.log
