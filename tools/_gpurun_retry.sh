#!/bin/bash
# usage: tools/_gpurun_retry.sh <timeout> <logfile> [--gpus N] -- cmd
to=$1; log=$2; shift 2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $to "$@" > $log 2>&1
  if grep -q "status=transient\|rc=3\b" $log && ! grep -q "charged=[1-9]" $log; then sleep 120; continue; fi
  break
done
