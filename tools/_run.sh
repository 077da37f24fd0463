timeout 1500 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "multigrid or verified_by_c_port" 2>&1 | tail -30 > gpurun_out/r02_mg_t1.log
cat gpurun_out/r02_mg_t1.log
timeout 1500 python tools/r02_probe_mg.py 256,128,128 1e-8 2:6 3:8 3:15 4:20 > gpurun_out/r02_mg_p2.log 2>&1
grep MG gpurun_out/r02_mg_p2.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n1_d.json'))
print(d['value'], d['e2e'], d['multigrid_run'], d['converged_run'])
"
