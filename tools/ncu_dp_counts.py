"""DP instructions and DRAM bytes per contract from ncu --set full captures -> profiles/dp_per_contract.json
(read by bench.py for the FP64 roofline, so that the figures follow the kernels instead of living in the source):

    python tools/ncu_dp_counts.py TAG workload=rep.ncu-rep:contracts [workload=repA+repB:contracts ...]

dp_inst  = DFMA+DMUL+DADD+DSETP+DMNMX executed on ACTIVE lanes / contracts (the algorithmic work),
dp_slots = warp-level DP issues x 32 / contracts (what occupies the FP64 pipe, inactive lanes included),
traffic_bytes = (dram__bytes_read.sum + dram__bytes_write.sum) / contracts."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "dp_per_contract.json")
DP = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")


def counts(rep, contracts):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    ix = {n: i for i, n in enumerate(h)}
    warp = thr = tot = 0
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]].strip())
        n = int(r[ix["Instructions Executed"]])
        tot += n
        if m and m.group(1) in DP:
            warp += n
            thr += int(r[ix["Thread Instructions Executed"]])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hh, units, vals = rr[0], rr[1], rr[2]

    def metric(name):
        i = hh.index(name)
        v = float(vals[i].replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[i], 1.0)
    traffic = metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")
    return {"dp_inst": round(thr / contracts, 1), "dp_slots": round(warp * 32 / contracts, 1), "inst_slots": round(tot * 32 / contracts, 1),
            "traffic_bytes": round(traffic / contracts, 1), "contracts": contracts, "kernel": vals[hh.index("Kernel Name")][:60],
            "fp64_pipe_pct": float(vals[hh.index("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")]),
            "issue_active_pct": float(vals[hh.index("smsp__issue_active.avg.pct_of_peak_sustained_active")]),
            "duration_us": round(float(vals[hh.index("gpu__time_duration.sum")].replace(",", ""))
                                 * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[hh.index("gpu__time_duration.sum")], 1.0), 3),
            "report": os.path.basename(rep)}


if __name__ == "__main__":
    tag = sys.argv[1]
    try:
        table = json.load(open(OUT))
    except Exception:
        table = {"kernels": {}}
    for a in sys.argv[2:]:
        wl, rest = a.split("=")
        rep, n = rest.rsplit(":", 1)
        parts = [counts(x, int(n)) for x in rep.split("+")]  # a workload of several launches per step: repA+repB
        tot = dict(parts[0])
        for q in parts[1:]:
            for k in ("dp_inst", "dp_slots", "inst_slots", "traffic_bytes", "duration_us"):
                tot[k] = round(tot[k] + q[k], 1)
            tot["kernel"] += " + " + q["kernel"]
            tot["report"] += " + " + q["report"]
        if len(parts) > 1:
            for k in ("fp64_pipe_pct", "issue_active_pct"):  # time-weighted over the launches of one step
                tot[k] = round(sum(q[k] * q["duration_us"] for q in parts) / sum(q["duration_us"] for q in parts), 2)
        table["kernels"][wl] = tot
        table["kernels"][wl]["capture"] = tag
        print(wl, table["kernels"][wl])
    table["source"] = f"ncu --set full captures, tools/ncu_dp_counts.py; per-kernel `capture` tags name the profiles/ summary"
    if "surface" not in table["kernels"] and "greeks" in table["kernels"] and "price" in table["kernels"]:
        pass
    json.dump(table, open(OUT, "w"), indent=1)
