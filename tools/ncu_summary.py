"""Summarises an .ncu-rep (raw + source pages) into text: python tools/ncu_summary.py rep [launch_idx]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("---- launch", r[hdr.index("ID")] if "ID" in hdr else "")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w} = {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split("Kernel Name")
for blk in blocks[1:2 if len(sys.argv) < 3 else None]:
    rr = list(csv.reader(("Kernel Name" + blk).splitlines()))
    h = rr[1]
    ix = {x: i for i, x in enumerate(h)}
    data = [r for r in rr[2:] if len(r) == len(h)]
    ti = sum(int(r[ix["Instructions Executed"]]) for r in data)
    ts = sum(int(r[ix["# Samples"]]) for r in data) or 1
    mix, smp = collections.Counter(), collections.Counter()
    for r in data:
        parts = r[ix["Source"]].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.split(".")[0]
        mix[op] += int(r[ix["Instructions Executed"]])
        smp[op] += int(r[ix["# Samples"]])
    print("  opcode mix (executed % / stall-sample %):")
    for op, c in mix.most_common(18):
        print(f"    {op:8s} {c / ti * 100:5.1f}  {smp[op] / ts * 100:5.1f}")
    print("  stall reasons (% of samples):")
    for x in h:
        if x.startswith("stall_") and "Not Issued" not in x:
            v = sum(int(r[ix[x]] or 0) for r in data)
            if v * 100 > ts:
                print(f"    {x:24s} {v / ts * 100:5.1f}")
