"""Executed warp instructions per SASS function of a kernel (the kernel body and every out-of-line device function it
calls), from an ncu report:   python tools/ncu_funcs.py rep.ncu-rep
Function boundaries = the targets of the CALL instructions; the first line of each function names its source line."""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, cur, curfile = None, None, None
inst = {}  # address -> (n, thr, dp, src)
DP = ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")
targets = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        curfile = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, n in enumerate(hdr):
            ix.setdefault(n, i)
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[0] != "":
        cur = f"{curfile}:{r[0]} {r[1].strip()[:70]}"
        continue
    if r[2] in ("...", "-"):
        continue
    try:
        n, t = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])
    except ValueError:
        continue
    a = int(r[2], 16)
    sass = r[3].strip()
    m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", sass)
    op = m.group(1) if m else ""
    if a not in inst:  # an instruction appears once per source line it is attributed to: count it once
        inst[a] = (n, t, n if op in DP else 0, cur, sass)
    if "CALL" in sass:
        mm = re.search(r"0x([0-9a-f]+)", sass)
        if mm:
            targets.add(int(mm.group(1), 16))
addrs = sorted(inst)
starts = sorted(t for t in targets if t in inst)
starts = [addrs[0]] + [s for s in starts if s != addrs[0]]
tot = sum(v[0] for v in inst.values())
print(f"{rep}: {tot} warp instructions")
for i, s in enumerate(starts):
    e = starts[i + 1] if i + 1 < len(starts) else addrs[-1] + 16
    sel = [inst[a] for a in addrs if s <= a < e]
    n = sum(v[0] for v in sel)
    t = sum(v[1] for v in sel)
    dp = sum(v[2] for v in sel)
    calls = inst[s][0]
    print(f"{s:#x} {len(sel):5d} sass  {n:11d} {n / tot * 100:5.1f}%  dp {dp / max(n, 1) * 100:4.1f}%  lanes {t / max(n, 1):4.1f}  entered {calls:8d}  {inst[s][3]}")
