#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer.py -m gpu -x -q > $O/pytest_peer16.log 2>&1; echo "pytest peer rc=$?"
tail -3 $O/pytest_peer16.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 3 --warmup 3 --no-c5 --no-cpu > $O/bench_2gpu_l.json 2> $O/bench_2gpu_l.err
python - <<'P'
import json
for line in open('gpurun_out/bench_2gpu_l.json'):
    if line.startswith('{'):
        d=json.loads(line); c=d['c4_summary']; print({k:round(v,3) for k,v in c.items() if k.endswith("_ms") or "kernel" in k}); print(c['parity_check']['ok'], c['parity_check']['rel_l2']); print([ (r['rank'], round(r['walk_kernel_ms'],2)) for r in c.get('per_rank',[])])
        t=d['tree_summary']; print('tree', round(t['ms_per_step'],3), round(t['build_ms'],3), round(t['walk_kernel_ms'],3), t['parity_check']['ok'])
P
tail -2 $O/bench_2gpu_l.err
