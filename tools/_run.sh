timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_tma -s 2 -c 1 -f -o gpurun_out/r02_tma_v1 python tools/profile_ring.py 256,128,128 9 > gpurun_out/ncu_tma1.log 2>&1
tail -2 gpurun_out/ncu_tma1.log
