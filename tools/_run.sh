timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_worker.py > gpurun_out/r02_mgpu8_fused.log 2>&1
grep "MGPU_OK\|AssertionError" gpurun_out/r02_mgpu8_fused.log | head -3
for f in 1 0; do
TOPOPT_CG_FUSED_MGPU=$f timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 --no-converged-run > gpurun_out/r02_bench_n8_fused$f.json 2> gpurun_out/r02_bench_n8_fused$f.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n8_fused$f.json'))
print('fused=$f', d['value'], d['e2e']['value'], d['cg_iteration']['ms'], d['parity_check']['ok'], d['config']['objective'])
" || tail -5 gpurun_out/r02_bench_n8_fused$f.err
done
