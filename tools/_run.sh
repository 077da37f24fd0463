cp topopt.jl_b200/libtopopt_cuda.so /tmp/good.so
for v in nst6 nosleep; do
  cp tools/probes/variants/$v.so topopt.jl_b200/libtopopt_cuda.so
  echo "== $v"
  timeout 300 python tools/r02_probe_fused.py 256,128,128 TOPOPT_CG_FUSED_TMA=1 TOPOPT_CG_FUSED_TMA=0 2>&1 | tail -2
done
cp /tmp/good.so topopt.jl_b200/libtopopt_cuda.so
