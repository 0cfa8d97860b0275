#!/bin/bash
# Runs on the GPU box (under gpurun): A/B of engine builds (build.py --variant NAME ... -> build/libpvv_NAME.so).
#   VARIANTS="base main foo" WL="greeks price iv iv_greeks" TAG=r02a bash tools/ab_perf.sh
# Per variant: the golden-vector parity tests (bit comparison with the reference's outputs), then the device-resident
# throughput of the kernels.  FUZZ=1 adds the extreme-input fuzz against the oracle for the first variant.
mkdir -p gpurun_out
TAG=${TAG:-ab}
OUT=gpurun_out/${TAG}_ab.txt
: > $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader >> $OUT
for V in ${VARIANTS:-main}; do
  if [ "$V" = "main" ]; then unset PVV_LIB; else export PVV_LIB=$PWD/py_vollib_vectorized_b200/build/libpvv_$V.so; fi
  echo "== variant $V" | tee -a $OUT
  python -m pytest ${PYTESTS:-tests/test_gpu_parity.py} -m gpu -x -q 2>&1 | tail -2 | tee -a $OUT
  for W in ${WL:-greeks price iv iv_greeks}; do
    N=""
    [ "$W" = "iv" ] && N="--n ${IV_N:-4000000}"
    [ "$W" = "iv_greeks" ] && N="--n ${IVG_N:-2000000}"
    python bench.py --workload $W --steps ${STEPS:-60} --warmup 10 --no-cpu-baseline $N 2> gpurun_out/${TAG}_$V_$W.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%-10s %.3f G/s  %.4f ms/step  e2e %.3f G/s  clk %s' % ('$W', d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['clocks']['sm_mhz'] if d['clocks'] else None))" | tee -a $OUT || tail -3 gpurun_out/${TAG}_$V_$W.err | tee -a $OUT
  done
done
unset PVV_LIB
if [ -n "$FUZZ" ]; then FUZZ_N=${FUZZ_N:-200000} python tools/fuzz_parity.py 2>&1 | tail -${FUZZ_TAIL:-8} | tee -a $OUT; fi
if [ -n "$ILP" ]; then python tools/probe_ilp.py 2>&1 | tee -a $OUT; fi
