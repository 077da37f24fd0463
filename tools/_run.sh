rm -f gpurun_out/r02_small_configs.jsonl
for c in 1 2 3 5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/r02_small_configs.jsonl; done
python - <<'P'
import json
for l in open('gpurun_out/r02_small_configs.jsonl'):
    d=json.loads(l)
    print(d['config']['workload'][:60], d['value'], d['config']['cg'], d.get('cg_iteration'), d['config']['objective'])
P
