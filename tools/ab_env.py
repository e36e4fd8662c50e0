"""A/B on ONE box of builds / settings of the library, each run in its own subprocess (settings are read once per process), several
alternations, medians.  Each run is bench.py's own timed loop: 5 warm-up + 20 timed steps of forward + inverse with per-kernel CUDA
events (the driver's setting), then 300 back-to-back steps (sustained).

    python tools/ab_env.py [ROOT=<other checkout>] [NAME=VALUE ...]

compares the current tree / default environment against the variant (another checkout of the repo -- e.g. a worktree of the previous
round -- and / or extra environment variables: PQMF_H4_NBUF=2, AB_FLAGS=<extra flag bits>, AB_NBAND=<n_band>)."""
import json
import os
import statistics
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, os, sys, time
sys.path.insert(0, os.environ["AB_ROOT"])
import torch
import pqmf_b200 as pq
m = int(os.environ.get("AB_NBAND", "16"))
b, t = 64, 1 << 20
x = (0.5 * torch.randn(b, 1, t, device="cuda")).clamp_(-1, 1)
mod = pq.PQMF(100, m).cuda()
mod._flags |= int(os.environ.get("AB_FLAGS", "0"))
for _ in range(5):
    y = mod(x); o = mod.inverse(y)
torch.cuda.synchronize()
steps = 20
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
for k in range(steps):
    ev[k][0].record(); y = mod(x); ev[k][1].record(); o = mod.inverse(y); ev[k][2].record()
torch.cuda.synchronize()
ta = sum(e[0].elapsed_time(e[1]) for e in ev) / steps
ts = sum(e[1].elapsed_time(e[2]) for e in ev) / steps
tot = ev[0][0].elapsed_time(ev[-1][2]) / steps
time.sleep(0.5)
n = int(os.environ.get("AB_STEPS", "300"))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): mod.inverse(mod(x))
e1.record(); torch.cuda.synchronize()
print(json.dumps({"ta": ta, "ts": ts, "step": tot, "sus": e0.elapsed_time(e1) / n}))
'''


def run(root, env_extra):
    env = dict(os.environ, AB_ROOT=root, **env_extra)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(r.stderr[-2000:])
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def main():
    extra = dict(a.split("=", 1) for a in sys.argv[1:])
    root = extra.pop("ROOT", HERE)
    rounds = int(os.environ.get("AB_ROUNDS", "4"))
    res = {"default": [], "variant": []}
    for _ in range(rounds):
        time.sleep(2.0)
        res["default"].append(run(HERE, {}))
        time.sleep(2.0)
        res["variant"].append(run(root, extra))
    ns = 64 * (1 << 20)
    for name, rs in res.items():
        ta, ts, step, sus = (statistics.median(r[k] for r in rs) for k in ("ta", "ts", "step", "sus"))
        tag = f"{os.path.basename(root) if root != HERE else ''} {extra}" if name == "variant" else ""
        print(f"{name:8s} {tag}: 20 steps: analysis {ta:.4f} synthesis {ts:.4f} ms/step {step:.4f} -> dominant {8 * ns / max(ta, ts) * 1e-6 / 6552.6:.3f}, round trip "
              f"{16 * ns / step * 1e-6 / 6552.6:.3f} | sustained {sus:.4f} ms/step = {16 * ns / sus * 1e-6 / 6552.6:.3f}   (all steps: {[round(r['step'], 4) for r in rs]}; "
              f"all sustained: {[round(r['sus'], 4) for r in rs]})", flush=True)


if __name__ == "__main__":
    main()
