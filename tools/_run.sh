timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 --no-converged-run > gpurun_out/r02_bench_n8_push.json 2> gpurun_out/r02_bench_n8_push.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n8_push.json'))
print('n8', d['value'], d['e2e']['value'], d['cg_iteration']['ms'], d['parity_check'], d['config']['objective'])
" || tail -5 gpurun_out/r02_bench_n8_push.err
