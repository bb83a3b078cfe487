"""Executed warp-instruction mix by SASS opcode from an .ncu-rep (ncu --set full --import-source on).
Usage: ncu_sass_mix.py report.ncu-rep kernel-regex [top]"""
import collections
import csv
import io
import subprocess
import sys

rep, pattern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name",
                      f"regex:{pattern}"], capture_output=True, text=True).stdout
mix, col, seen_kernel = collections.Counter(), None, 0
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1:
            break                                   # first matching launch only
        continue
    if r[0] == "Address":
        col = {h: i for i, h in enumerate(r)}
        continue
    if col is None or len(r) <= col["Instructions Executed"]:
        continue
    try:
        n = int(r[col["Instructions Executed"]])
    except ValueError:
        continue
    ops = r[col["Source"]].split()
    if not ops:
        continue
    op = ops[1] if ops[0].startswith("@") and len(ops) > 1 else ops[0]
    mix[op.split(".")[0]] += n
total = sum(mix.values()) or 1
print(f"# {pattern}: {total} warp instructions")
for op, n in mix.most_common(top):
    print(f"{100 * n / total:5.1f}%  {op}")
