#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests that gate every kernel change, then the device-resident
# throughput of the four kernels.  Usage: bash tools/quick_perf.sh [pytest -k expression]
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -3
for W in ${WL:-greeks price iv iv_greeks}; do
  python bench.py --workload $W --steps 100 --warmup 10 --no-cpu-baseline 2> gpurun_out/q_$W.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-10s %.3f G/s  %.3f ms/step  fp64 %.3f  e2e %.3f G/s  clk %s' % ('$W', d['value']/1e9, d['ms_per_step'], d['fp64']['frac'], d['e2e']['value']/1e9, d['clocks']['sm_mhz'] if d['clocks'] else None))" || tail -3 gpurun_out/q_$W.err
done
