"""Summarise an .ncu-rep (read on the CPU box): headline metrics per kernel + opcode mix + top stall sites.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--src]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", "sm__cycles_active.avg", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"]
for w in want:
    for h in hdr:
        if h == w or h.endswith(w):
            print(f"{h[-80:]:80s} " + " | ".join(r[idx[h]][:40] for r in data) + f"  {units[idx[h]]}")
            break
print("stalls per issued instruction:")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        vals = [float(r[idx[h]]) for r in data]
        if max(vals) >= 0.05:
            print(f"  {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} " + " | ".join(f"{v:6.2f}" for v in vals))
if "--src" in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    for b in blocks:
        if b["name"] in seen:
            continue
        seen.add(b["name"])
        h = b["rows"][0]
        ix = {k: i for i, k in enumerate(h)}
        d = b["rows"][1:]
        f = lambda r, k: float(r[ix[k]]) if r[ix[k]] not in ("", None) else 0.0
        tot = sum(f(r, "Instructions Executed") for r in d)
        print("=====", b["name"][:70], "warp instr", int(tot))
        c, s = Counter(), Counter()
        for r in d:
            t = r[ix["Source"]].split()
            if not t:
                continue
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            c[op] += f(r, "Instructions Executed")
            s[op] += f(r, "# Samples")
        for op, n in c.most_common(18):
            print(f"   {op:10s} {int(n):10d} {100*n/tot:5.1f}%  samples {int(s[op])}")
        print("   -- top stall sites")
        for r in sorted(d, key=lambda r: -f(r, "# Samples"))[:14]:
            st = {k.replace("stall_", ""): int(f(r, k)) for k in h if k.startswith("stall_") and "Not" not in k and f(r, k) > 0}
            print("    ", r[ix["Source"]][:56].ljust(56), int(f(r, "# Samples")), st)
