"""Per-kernel SASS opcode histogram of libpqmf_b200.so (cuobjdump -sass; runs on the CPU box).
    python tools/sass_ops.py [substring ...] > profiles/r02_sass_ops.txt
Lists, for every kernel whose demangled name contains one of the substrings (default: the Hankel kernels), the instruction count
and the opcode histogram -- the evidence for UTCHMMA (tcgen05.mma) / LDTM (tcgen05.ld) / UBLKCP (cp.async.bulk) / STG.256 etc."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pseudo-quadrature-mirror-filter_b200", "libpqmf_b200.so")


def main():
    want = sys.argv[1:] or ["h4_analysis_kernel<(int)16, (bool)1", "h4_synthesis_kernel<(int)16, (bool)1", "h4_analysis_stream", "h4_synthesis_stream", "f16_"]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = {}
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    mangled = [f.split("\n", 1)[0].strip() for f in funcs]
    dem = subprocess.run(["cu++filt"] + mangled, capture_output=True, text=True).stdout.splitlines()
    for m, d, body in zip(mangled, dem, funcs):
        if not any(w in d for w in want):
            continue
        ops = collections.Counter()
        full = collections.Counter()
        for line in body.splitlines():
            mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
            if mm:
                ops[mm.group(1)] += 1
                if mm.group(1) in ("UTCHMMA", "LDTM", "UBLKCP", "STG", "LDG", "STS", "LDS", "SYNCS", "UTCBAR", "F2FP", "FFMA2", "FMUL2", "SHFL"):
                    full[mm.group(1) + mm.group(2)] += 1
        total = sum(ops.values())
        name = d[:d.index(">(") + 1] if ">(" in d else d.split("(")[0]
        print(f"== {name}   [{total} instructions]")
        print("   " + "  ".join(f"{k}:{v}" for k, v in ops.most_common(28)))
        print("   detail: " + "  ".join(f"{k}:{v}" for k, v in sorted(full.items())))
    return 0


if __name__ == "__main__":
    sys.exit(main())
