#!/bin/bash
O=gpurun_out
NP=${1:-4}
B200_BENCH_WALK_MATRIX=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $NP --steps 3 --warmup 3 --no-c5 --no-cpu > $O/bench_${NP}gpu_k.json 2> $O/bench_${NP}gpu_k.err
python - <<P
import json
for line in open('gpurun_out/bench_${NP}gpu_k.json'):
    if line.startswith('{'):
        c=json.loads(line)['c4_summary']; print({k:round(v,3) for k,v in c.items() if k.endswith("_ms") or "kernel" in k}); print(c.get("per_rank"))
P
tail -2 $O/bench_${NP}gpu_k.err
