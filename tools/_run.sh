timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_fused -s 2 -c 1 -f -o gpurun_out/r02_fused_final python tools/profile_ring.py 256,128,128 9 > gpurun_out/ncu_fused_final.log 2>&1
tail -2 gpurun_out/ncu_fused_final.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-converged-run > gpurun_out/r02_bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_launches_bench.csv
TOPOPT_SKIP_FULL_SIZE=1 timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "one_kernel_cg_iteration_matches or multigrid" > gpurun_out/r02_memcheck_fused.log 2>&1
tail -5 gpurun_out/r02_memcheck_fused.log
TOPOPT_SKIP_FULL_SIZE=1 timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "dense_rhs and nels1" > gpurun_out/r02_racecheck_fused.log 2>&1
tail -8 gpurun_out/r02_racecheck_fused.log
