timeout 1200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "dense_rhs or one_kernel" 2>&1 | tail -30 > gpurun_out/r02_fused_t7.log
cat gpurun_out/r02_fused_t7.log
timeout 900 python tools/r02_probe_fused.py 256,128,128 TOPOPT_CG_FUSED=1 TOPOPT_CG_FUSED=0 > gpurun_out/r02_fused_p10.log 2>&1
cat gpurun_out/r02_fused_p10.log
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n1_c.json'))
print(d['value'], d['config']['cg'], d['reference_recurrence'], d['converged_run'])
"
