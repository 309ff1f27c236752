"""List every mbarrier wait loop (SYNCS.PHASECHK.TRYWAIT) of the kernels in an .ncu-rep with its stall samples,
so that the inlined mbar_wait call sites can be told apart by execution count and neighbourhood.

usage: python tools/ncu_waits.py report.ncu-rep
"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
ks, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}
        ks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r and r[0].startswith("0x"):
        cur["rows"].append(r)
for k in ks:
    h, R = k["hdr"], k["rows"]
    ia, ie, isrc = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed"), h.index("Source")
    tot = sum(int(r[ia]) for r in R)
    print("=" * 100)
    print(k["name"][:100], "instructions:", len(R), "samples:", tot)
    for i, r in enumerate(R):
        if "TRYWAIT" in r[isrc]:
            j = i
            while j < len(R) and j < i + 12 and "BRA" not in R[j][isrc]:
                j += 1
            s = sum(int(x[ia]) for x in R[i:j + 1])
            ctx = " | ".join(x[isrc].split()[0] if not x[isrc].startswith("@") else x[isrc].split()[1] for x in R[max(0, i - 4):i])
            print(f"  sass#{i:5d} first-poll exec={int(r[ie]):>9} loop samples={s:6d} ({100 * s / tot:4.1f}%)   before: {ctx}")
