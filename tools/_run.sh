timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_gputest6.log
cat gpurun_out/r02_gputest6.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_g.json 2> gpurun_out/r02_bench_n1_g.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_n1_g.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['clocks'], d['reference_recurrence']['objective_rel_diff_vs_headline'])
"
