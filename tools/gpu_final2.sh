#!/bin/bash
# round-2 final: Barnes-Hut as the main bench line (BASELINE config 3), uniform and Zel'dovich inputs
O=gpurun_out
python bench.py --workload tree --steps 10 > $O/bench_final_tree.json 2> $O/bench_final_tree.err; echo "tree rc=$?"
python bench.py --workload tree --ic zeldovich --steps 10 --no-cpu > $O/bench_final_tree_zel.json 2> $O/bench_final_tree_zel.err; echo "tree zel rc=$?"
python - <<'P'
import json
for f in ('bench_final_tree','bench_final_tree_zel'):
    for line in open(f'gpurun_out/{f}.json'):
        if line.startswith('{'):
            d=json.loads(line); print(f, d['metric'], d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'), d.get('parity_check',{}).get('ok'), d.get('cpu_baseline',{}).get('value'))
P
