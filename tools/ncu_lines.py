"""Executed warp instructions and stall samples per CUDA source line (and per opcode class within a line range)
from an ncu report captured with --import-source on:
   python tools/ncu_lines.py rep.ncu-rep [top=60] [file-substring:lo-hi ...]
With file:lo-hi arguments prints the totals of those line ranges (e.g. pvv_math.cuh:461-537 = cody_erfc)."""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 60
ranges = [a for a in sys.argv[2:] if ":" in a]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
DP = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
cur_file, hdr = None, None
lines = {}  # (file, line) -> dict
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, n in enumerate(hdr):
            ix.setdefault(n, i)
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[0] not in ("", "0"):
        cur = lines.setdefault((cur_file, int(r[0])), dict(src=r[1].strip(), n=0, dp=0, thr=0, samples=0, ops=collections.Counter(), st=collections.Counter()))
        continue
    if r[0] == "0":
        cur = lines.setdefault((cur_file, 0), dict(src="", n=0, dp=0, thr=0, samples=0, ops=collections.Counter(), st=collections.Counter()))
        continue
    if cur is None or r[2] in ("...", "-"):
        continue
    sass = r[3].strip()
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    op = m.group(1) if m else sass[:10]
    base = op.split(".")[0]
    if base == "IMAD" and ".MOV" in op:
        base = "IMAD.MOV"
    try:
        n = int(r[ix["Instructions Executed"]])
        t = int(r[ix["Thread Instructions Executed"]])
        s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    cur["n"] += n
    cur["thr"] += t
    cur["samples"] += s
    cur["ops"][base] += n
    if base in DP:
        cur["dp"] += n
    for k in ("stall_wait", "stall_barrier", "stall_math", "stall_no_inst", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "stall_not_selected", "stall_dispatch"):
        if k in ix:
            try:
                cur["st"][k] += int(r[ix[k]])
            except ValueError:
                pass
tot = sum(v["n"] for v in lines.values())
ts = sum(v["samples"] for v in lines.values())
print(f"{rep}: {tot} warp instructions, {ts} samples")
if ranges:
    for a in ranges:
        f, lh = a.split(":")
        lo, hi = (int(x) for x in lh.split("-"))
        sel = [v for (ff, ll), v in lines.items() if f in ff and lo <= ll <= hi]
        n = sum(v["n"] for v in sel)
        dp = sum(v["dp"] for v in sel)
        thr = sum(v["thr"] for v in sel)
        s = sum(v["samples"] for v in sel)
        ops = collections.Counter()
        st = collections.Counter()
        for v in sel:
            ops.update(v["ops"])
            st.update(v["st"])
        print(f"{a:34s} instr {n:11d} {n / tot * 100:5.1f}%  DP share {dp / max(n, 1) * 100:5.1f}%  lanes {thr / max(n, 1):5.1f}  samples {s / max(ts, 1) * 100:5.1f}%")
        print("     ops: " + " ".join(f"{k}:{v / max(n, 1) * 100:.0f}%" for k, v in ops.most_common(12)))
        print("     stalls: " + " ".join(f"{k[6:]}:{v / max(s, 1) * 100:.0f}%" for k, v in st.most_common(6)))
else:
    for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1]["n"])[:top]:
        print(f"{f}:{l:<5d} {v['n']:10d} {v['n'] / tot * 100:5.1f}% dp {v['dp'] / max(v['n'], 1) * 100:3.0f}% lanes {v['thr'] / max(v['n'], 1):4.1f} smp {v['samples'] / max(ts, 1) * 100:4.1f}%  {v['src'][:90]}")
