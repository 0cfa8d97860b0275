#!/bin/bash
# Runs on the GPU box (under gpurun): the parity tests that gate every kernel change, then the device-resident
# throughput of the kernels — for the in-tree engine, or for A/B builds (build.py --variant NAME ...):
#   VARIANTS="base foo" WL="greeks iv" bash tools/quick_perf.sh
mkdir -p gpurun_out
for V in ${VARIANTS:-main}; do
  if [ "$V" = "main" ]; then unset PVV_LIB; else export PVV_LIB=$PWD/py_vollib_vectorized_b200/build/libpvv_$V.so; fi
  echo "== variant $V"
  python -m pytest tests -m gpu -x -q 2>&1 | tail -1
  for W in ${WL:-greeks price iv iv_greeks}; do
    python bench.py --workload $W --steps ${STEPS:-100} --warmup 10 --no-cpu-baseline 2> gpurun_out/q_$W.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-10s %.3f G/s  %.4f ms/step  e2e %.3f G/s  clk %s' % ('$W', d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['clocks']['sm_mhz'] if d['clocks'] else None))" || tail -3 gpurun_out/q_$W.err
  done
done
