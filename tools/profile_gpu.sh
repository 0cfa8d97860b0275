#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the default bench + one ncu --set full capture per kernel.
# Usage: bash tools/profile_gpu.sh <tag>
TAG=${1:-rXX}
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/${TAG}_plain_greeks.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv $B > gpurun_out/${TAG}_ncu_launches.log 2>&1
for W in ${WL:-greeks price iv iv_greeks}; do
  N=""; K="${W}_kernel"
  [ "$W" = "iv" ] && N="--n 3000000" && K="^iv_kernel"
  [ "$W" = "iv_greeks" ] && N="--n 2000000"
  [ "$W" = "greeks" ] && K="^greeks_kernel"
  C="python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline $N"
  $C > gpurun_out/${TAG}_plain_$W.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/${TAG}_prof_$W -f $C > gpurun_out/${TAG}_ncu_$W.log 2>&1
done
ls -la gpurun_out | grep $TAG
