"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel the launch count, total and mean device time and
the share of all listed GPU time, with the whole-batch launches of the device-timed loop (>= 60 us) apart from the row-chunk launches
of the host-buffer e2e path.    python tools/launch_summary.py gpurun_out/r02_bench_launches.csv > profiles/r02_bench_launches_summary.txt"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                ns = float(d["Metric Value"].replace(",", ""))
                name = d["Kernel Name"].split("(")[0].replace("void ", "")[:70]
                cls = "whole batch" if ns >= 60e3 else "row chunk / small"
                agg.setdefault((name, cls), []).append(ns)
    tot = sum(sum(v) for v in agg.values())
    print(f"# {sys.argv[1]}: {sum(len(v) for v in agg.values())} launches, {tot / 1e6:.3f} ms of GPU time (cold-cache, serialised under ncu: compare shares)")
    print(f"{'kernel':72s} {'class':18s} {'n':>4s} {'total us':>10s} {'mean us':>9s} {'share':>6s}")
    for (name, cls), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{name:72s} {cls:18s} {len(v):4d} {sum(v) / 1e3:10.1f} {sum(v) / len(v) / 1e3:9.1f} {sum(v) / tot:6.3f}")
    ours = sum(sum(v) for (n, _), v in agg.items() if n.startswith("pqmf::"))
    print(f"pqmf:: kernels: {ours / tot:.3f} of the listed GPU time; the rest is torch's input generation / checks outside the timed region")


if __name__ == "__main__":
    main()
