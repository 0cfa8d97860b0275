"""Executed warp instructions by opcode (and stall samples) from an ncu report's source page.
   python tools/ncu_opcodes.py rep.ncu-rep [contracts]"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
contracts = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
ops, samples, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) < len(h):
        continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(1) if m else src[:10]
    base = op.split(".")[0]
    if base == "IMAD" and ".MOV" in op:
        base = "IMAD.MOV"
    if base == "MUFU":
        base = op
    n = int(r[ix["Instructions Executed"]])
    ops[base] += n
    thr[base] += int(r[ix["Thread Instructions Executed"]])
    samples[base] += int(r[ix["# Samples"]])
tot, ts = sum(ops.values()), sum(samples.values())
dp = sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(f"{rep}: warp instructions {tot}, DP {dp} ({dp / tot * 100:.1f}%), active threads/instr {sum(thr.values()) / tot:.2f}")
if contracts:
    dpt = sum(thr[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
    print(f"per contract: {tot * 32 / contracts:.0f} thread-instruction slots, {dp * 32 / contracts:.0f} DP slots (warp-level x 32: what occupies the pipe), "
          f"{dpt / contracts:.0f} DP instructions on active lanes (the algorithmic work)")
for k, v in ops.most_common(24):
    print(f"  {k:14s} {v:12d} {v / tot * 100:5.1f}%   stall samples {samples[k] / max(ts, 1) * 100:5.1f}%")
