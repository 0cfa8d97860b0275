"""Executed CALL instructions by call site (out-of-line evaluators, the IEEE dividers' slow paths, rare branches of exp/log)
from an ncu report captured with --import-source on:   python tools/ncu_calls.py rep.ncu-rep [top=40]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr, cur, curfile, calls, total = None, None, None, {}, 0
for r in csv.reader(out.splitlines()):
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
        cur = (curfile, r[0], r[1].strip()[:90])
        continue
    try:
        n, t = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])
    except ValueError:
        continue
    if "CALL" in r[3] and n:
        calls[r[2]] = (n, t, r[3].strip().split()[-1][-6:], cur)
for addr, (n, t, target, cur) in sorted(calls.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n:9d} calls at {t / n:4.1f} lanes  -> {target}  {cur[0]}:{cur[1]}  {cur[2]}")
