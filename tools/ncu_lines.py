"""Aggregate ncu stall samples / executed instructions per CUDA source line.
usage: python tools/ncu_lines.py rep.ncu-rep <kernel-regex> [N]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 28
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
def num(s):
    try: return float(s)
    except Exception: return 0.0
cur_file, ix, agg = None, None, {}
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": ix = {k: i for i, k in enumerate(r)}; continue
    if r[0] != "" and ix:
        try: line = int(r[0])
        except Exception: continue
        key = (cur_file, line, r[1].strip()[:100])
        a = agg.setdefault(key, [0.0, 0.0])
        a[0] += num(r[ix["# Samples"]]); a[1] += num(r[ix["Instructions Executed"]])
tot = sum(a[0] for a in agg.values()) or 1; toti = sum(a[1] for a in agg.values()) or 1
print("total samples", tot, "warp instr", toti)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]:18s}:{k[1]:4d} smp {a[0]:6.0f} ({100*a[0]/tot:4.1f}%) ins {a[1]:9.0f} ({100*a[1]/toti:4.1f}%)  {k[2]}")
