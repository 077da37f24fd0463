timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_worker.py > gpurun_out/r02_mgpu2_push.log 2>&1
grep "MGPU_OK\|AssertionError\|Error" gpurun_out/r02_mgpu2_push.log | head -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --no-converged-run > gpurun_out/r02_bench_n2_push.json 2> gpurun_out/r02_bench_n2_push.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n2_push.json'))
print('n2', d['value'], d['e2e']['value'], d['cg_iteration']['ms'], d['parity_check']['ok'], d['config']['objective'])
" || tail -5 gpurun_out/r02_bench_n2_push.err
