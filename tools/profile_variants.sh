#!/bin/bash
# Runs on the GPU box (under gpurun): one ncu --set full capture per (variant, workload).
#   VARIANTS="base nopair" WL="greeks iv" TAG=r02b bash tools/profile_variants.sh
mkdir -p gpurun_out
TAG=${TAG:-prof}
for V in ${VARIANTS:-main}; do
  if [ "$V" = "main" ]; then unset PVV_LIB; else export PVV_LIB=$PWD/py_vollib_vectorized_b200/build/libpvv_$V.so; fi
  for W in ${WL:-greeks iv}; do
    N=""; K="${W}_kernel"
    [ "$W" = "iv" ] && N="--n 3000000" && K="^iv_kernel"
    [ "$W" = "iv_greeks" ] && N="--n 2000000"
    [ "$W" = "greeks" ] && K="^greeks_kernel"
    [ "$W" = "analytic" ] && N="--n 4000000" && K="analytic_kernel_tma"
    C="python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline --no-e2e $N"
    PARTS=""
    [ "$W" = "iv_greeks" ] && [ "${PVV_IVG_FUSED:-0}" != "1" ] && PARTS="iv greeks"
    [ "$W" = "surface" ] && PARTS="price greeks"
    if [ -n "$PARTS" ]; then
      # two launches per step (IV -> greeks: K3 then K2; cfg5 surface: K1 then K2): one capture of each on the workload's own chain
      for P in $PARTS; do
        ncu --set full --clock-control none --import-source on -k regex:^${P}_kernel -s 3 -c 1 -o gpurun_out/${TAG}_${V}_${W}_$P -f $C > gpurun_out/${TAG}_${V}_${W}_${P}_ncu.log 2>&1
        tail -2 gpurun_out/${TAG}_${V}_${W}_${P}_ncu.log
      done
      continue
    fi
    ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/${TAG}_${V}_$W -f $C > gpurun_out/${TAG}_${V}_${W}_ncu.log 2>&1
    tail -2 gpurun_out/${TAG}_${V}_${W}_ncu.log
  done
done
ls -la gpurun_out | grep $TAG
