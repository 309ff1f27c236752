"""Writes profiles/r2_ncu_traffic.json from `ncu --set full` reports: per kernel, DRAM bytes moved per launch
(dram__bytes_read.sum + dram__bytes_write.sum), duration and the tensor / issue utilisation, tagged with the workload shape the
capture ran.  bench.py reads `roofline.traffic` from this file (and refuses it when kernel name or shape do not match).
  python tools/ncu_traffic.py <report.ncu-rep>=<shape tag> [...]        (runs here: ncu -i needs no GPU)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "sm__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main():
    out_path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    entries = json.load(open(out_path)) if os.path.exists(out_path) else []
    for arg in sys.argv[1:]:
        rep, shape = arg.split("=", 1)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            e = {"kernel": d["Kernel Name"], "shape": shape, "report": os.path.basename(rep)}
            for k in WANT:
                if k in d and d[k] != "":
                    v = float(d[k].replace(",", ""))
                    if k.startswith("dram__bytes"):
                        v *= UNIT.get(u[k], 1.0)
                    e[k] = v
                    if k == "gpu__time_duration.sum":
                        e["duration_unit"] = u[k]
            e["dram_bytes"] = e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
            entries = [x for x in entries if not (x["kernel"] == e["kernel"] and x["shape"] == e["shape"])] + [e]
    with open(out_path, "w") as f:
        json.dump(entries, f, indent=1)
    print(f"{out_path}: {len(entries)} entries")


if __name__ == "__main__":
    main()
