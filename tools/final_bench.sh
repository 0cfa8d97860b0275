#!/bin/bash
# Runs on the GPU box (under gpurun): GPU test tier, then one bench.py line per workload (default steps, CPU
# baseline included) and the reference arm, collected in gpurun_out/<tag>_bench_lines.jsonl.
TAG=${1:-rXX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/${TAG}_pytest_gpu.txt
: > gpurun_out/${TAG}_bench_lines.jsonl
for W in greeks price iv iv_greeks analytic surface; do
  python bench.py --workload $W 2> gpurun_out/${TAG}_bench_$W.err | tail -1 >> gpurun_out/${TAG}_bench_lines.jsonl
done
python bench.py --impl reference --steps 20 --warmup 3 2>> gpurun_out/${TAG}_bench_ref.err | tail -1 >> gpurun_out/${TAG}_bench_lines.jsonl
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench_lines.jsonl"):
    d = json.loads(l)
    print("%-40s %10.3f M/s  e2e %8.3f M/s  fp64 %s  hbm %s  cpu %s" % (d["metric"] + (" [ref]" if d.get("impl") else ""), d["value"] / 1e6, d["e2e"]["value"] / 1e6,
          ("%.3f" % d["fp64"]["frac"]) if "fp64" in d else "-", ("%.3f" % d["roofline"]["frac"]) if "roofline" in d else "-",
          ("%.1f M/s" % (d["cpu_baseline"]["value"] / 1e6)) if "cpu_baseline" in d else "-"))
PY
