"""Summarise an .ncu-rep (raw + source pages) into a short text file for profiles/."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("=" * 100)
    for h, u, v in zip(hdr, units, r):
        if any(h == k or h.startswith(k) for k in KEYS) and ("pct" in h or "." in h or h in KEYS):
            if any(h == k for k in KEYS) or h.startswith("sm__inst_executed_pipe_tensor"):
                print(f"{h:90s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
try:
    h = srows[1]
    ia, isrc, iex = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
    data = [r for r in srows[2:] if len(r) > max(ia, isrc, iex) and r[ia].isdigit()]
    tot = sum(int(r[ia]) for r in data)
    print("\ntop stall sites (warp stall samples, all):", tot, "samples")
    for r in sorted(data, key=lambda r: -int(r[ia]))[:25]:
        print(f"{int(r[ia]):7d} {100 * int(r[ia]) / tot:5.1f}%  exec={r[iex]:>9}  {r[isrc][:100]}")
    grp = collections.Counter()
    for r in data:
        grp[int(r[iex])] += int(r[ia])
    print("\nsamples grouped by per-instruction execution count (warp roles):")
    for k, v in sorted(grp.items(), key=lambda kv: -kv[1])[:8]:
        print(f"  exec={k:>9}  samples={v}  ({100 * v / tot:.1f}%)  instrs={sum(1 for r in data if int(r[iex]) == k)}")
except Exception as e:  # noqa: BLE001
    print("no source page:", e)
