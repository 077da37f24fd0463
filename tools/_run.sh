timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r02_gputest5.log
cat gpurun_out/r02_gputest5.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n1_f.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['kxu_alone']['ms_per_launch'], d['clocks'], d['multigrid_run']['step_s'], d['gpu_launches'])
"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
