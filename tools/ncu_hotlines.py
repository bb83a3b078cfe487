"""Per-CUDA-source-line share of executed warp instructions and stall samples from an .ncu-rep captured with
`--set full --import-source on` (kernels built with -lineinfo).
Usage: ncu_hotlines.py report.ncu-rep kernel-regex [top]"""
import csv
import io
import subprocess
import sys

rep, pattern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{pattern}"], capture_output=True, text=True).stdout
lines, fname, col = [], "?", None
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        col = {h: i for i, h in enumerate(r)}          # first "Source" is the CUDA text (dict keeps the LAST index)
    elif col and r[0].isdigit():
        try:
            lines.append((int(r[col["# Samples"]]), int(r[col["Instructions Executed"]]), f"{fname}:{r[0]}", r[1]))
        except ValueError:
            pass
ts = sum(v[0] for v in lines) or 1
ti = sum(v[1] for v in lines) or 1
print(f"# {pattern}: {ts} stall samples, {ti} warp instructions")
for s, i, where, text in sorted(lines, reverse=True)[:top]:
    print(f"{100 * i / ti:5.1f}% instr {100 * s / ts:5.1f}% stalls  {where}  {text.strip()[:150]}")
