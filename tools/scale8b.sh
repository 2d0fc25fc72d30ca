#!/bin/bash
# 8-GPU box: the 8-GPU bench line and the C++ sharded-simulation test after a change of the exchange
O=gpurun_out
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 \
     bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2_bench_8gpu.json 2> $O/r2_bench_8gpu.err; echo "bench8 rc=$?"
tests/host/_bin/shard_test > $O/r2_shard_test_8gpu_box.log 2>&1; echo "shard_test rc=$?"
tail -2 $O/r2_shard_test_8gpu_box.log
python - <<'P'
import json
for line in open('gpurun_out/r2_bench_8gpu.json'):
    if line.startswith('{'):
        d=json.loads(line); c=d['c4_summary']
        print('direct', d['value'], d['parity_check']['ok'])
        print({k:round(v,3) for k,v in c.items() if k.endswith("_ms") or "kernel" in k}, c['parity_check']['ok'])
        t=d['tree_summary']; print('tree', round(t['ms_per_step'],3), t['parity_check']['ok'])
P
