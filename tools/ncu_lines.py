"""Per-source-line warp-stall summary of the kernels in an .ncu-rep (needs -lineinfo + --import-source on).

usage: python tools/ncu_lines.py report.ncu-rep [top-n]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
kern, hdr, lines = None, None, []


def flush():
    if not kern or not lines:
        return
    tot = sum(l[2] for l in lines)
    print("=" * 110)
    print(kern[:140], " samples:", tot)
    for f, ln, smp, ex, src in sorted(lines, key=lambda l: -l[2])[:topn]:
        print(f"{smp:7d} {100 * smp / max(tot, 1):5.1f}%  exec={ex:>10}  {f}:{ln:<5} {src[:90]}")


fpath = ""
for row in csv.reader(out.splitlines()):
    if not row:
        continue
    if row[0] == "File Path":
        fpath = row[1].split("/")[-1]
        continue
    if row[0] == "Function Name":
        if row[1] != kern:
            flush()
            kern, lines = row[1], []
        continue
    if row[0] == "Line No":
        hdr = row
        i_s = hdr.index("Warp Stall Sampling (All Samples)")
        i_e = hdr.index("Instructions Executed")
        continue
    if hdr and row[0].isdigit():  # a CUDA source line (its SASS rows follow with an empty first column)
        try:
            lines.append((fpath, int(row[0]), int(row[i_s]), int(row[i_e]), row[1].strip()))
        except ValueError:
            pass
flush()
