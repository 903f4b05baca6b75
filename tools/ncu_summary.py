"""Extract the judged metrics from an .ncu-rep (read on the CPU box): per captured launch duration, DRAM
bytes, tensor-pipe %, occupancy, registers and the top warp-stall reasons."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum"]
for r in rows[2:]:
    print("=" * 100)
    for k in keys:
        if k in idx:
            print(f"{k:75s} {r[idx[k]]} {units[idx[k]]}")
    stalls = []
    for h in hdr:
        if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
            try:
                stalls.append((float(r[idx[h]].replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1
    print("stall samples: " + ", ".join(f"{n} {100*v/tot:.0f}%" for v, n in sorted(stalls, reverse=True)[:6]))

# optional: --traffic-json FILE  -> {kernel base name: mean dram bytes (read + write) per captured launch}
if "--traffic-json" in sys.argv:
    import json
    import os
    import re
    out = sys.argv[sys.argv.index("--traffic-json") + 1]
    acc = {}
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
        name = re.sub(r"<.*", "", re.sub(r"<unnamed>::|\(anonymous namespace\)::|void ", "", name))
        try:
            rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * mult.get(units[idx["dram__bytes_read.sum"]], 1.0)
            wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * mult.get(units[idx["dram__bytes_write.sum"]], 1.0)
        except (KeyError, ValueError):
            continue
        acc.setdefault(name, []).append(rd + wr)
    cur = {}
    if os.path.isfile(out):
        with open(out) as f:
            cur = json.load(f)
    for k, v in acc.items():
        cur[k] = sum(v) / len(v)
    with open(out, "w") as f:
        json.dump(cur, f, indent=1, sort_keys=True)
