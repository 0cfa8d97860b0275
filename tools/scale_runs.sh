#!/bin/bash
# Runs on a multi-GPU box (gpurun --gpus N): the N > 1 correctness test, then weak scaling (the driver's command) and
# the strong-scaled 100 M-contract cfg4 chain at every rank count up to the GPUs present.
#   TAG=r02 bash tools/scale_runs.sh
mkdir -p gpurun_out
TAG=${TAG:-scale}
G=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
nproc >> gpurun_out/${TAG}_topo.txt; free -g | head -2 >> gpurun_out/${TAG}_topo.txt
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_multi.txt
PORT=29500
for N in ${NS:-2 4 8}; do
  [ $N -le $G ] || continue
  PORT=$((PORT+1))
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT"
  $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_weak_$N.json 2> gpurun_out/${TAG}_weak_$N.err
  tail -c 400 gpurun_out/${TAG}_weak_$N.json | head -c 400; echo
  PORT=$((PORT+1))
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT"
  $TR bench.py --gpus $N --workload iv_greeks --scaling strong --steps 10 --warmup 3 > gpurun_out/${TAG}_strong_$N.json 2> gpurun_out/${TAG}_strong_$N.err
  tail -c 300 gpurun_out/${TAG}_strong_$N.err
done
python - <<'PY'
import json, glob, os
tag = os.environ.get("TAG", "scale")
for f in sorted(glob.glob(f"gpurun_out/{tag}_*_[0-9].json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    e = d.get("e2e", {})
    c = e.get("ceiling", {})
    print(f"{os.path.basename(f):28s} n_gpus {d['n_gpus']} {d['scaling']:6s} value {d['value']/1e9:7.3f} G/s  e2e {e.get('value',0)/1e9:6.3f} G/s  ceiling {c.get('value',0)/1e9:6.3f} G/s ({c.get('gbs_both_directions',0):5.1f} GB/s)  parity {d.get('parity_checked',{}).get('ok')} gather_ms {d.get('gather_ms')} gather_ok {d.get('gather_ok')}")
    for k, v in (d.get("kernels") or {}).items():
        print(f"    {k:10s} value {v['value']/1e9:7.3f} G/s  e2e {v.get('e2e',{}).get('value',0)/1e9:6.3f}  ceiling {v.get('e2e',{}).get('ceiling',{}).get('value',0)/1e9:6.3f} parity {v.get('parity_checked',{}).get('ok')}")
PY
