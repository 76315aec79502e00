#!/usr/bin/env python3
"""cuobjdump -sass summary of the plane-streaming kernels of the built device library: instruction count and the histogram of the
Blackwell-relevant mnemonics per hot kernel, plus the TMA issue excerpt of the fused Chebyshev kernel.
Usage: python tools/sass_summary.py > profiles/rNN_sass_hot_kernels.txt"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dealii_spirk_b200", "libspirk_b200.so")
KERNELS = [
    ("fused Chebyshev step (own diagonal), NPT=2", "_ZN5spirk4k_v3ILi4ELi8ELi8ELi3ELi2ELi1EEEvNS_6V3ArgsE"),
    ("plain apply (stage vmult), NPT=2", "_ZN5spirk4k_v3ILi4ELi8ELi8ELi0ELi2ELi1EEEvNS_6V3ArgsE"),
    ("residual, NPT=2", "_ZN5spirk4k_v3ILi4ELi8ELi8ELi1ELi2ELi1EEEvNS_6V3ArgsE"),
    ("fused first two Chebyshev iterations, NPT=2", "_ZN5spirk4k_v3ILi4ELi8ELi8ELi4ELi2ELi1EEEvNS_6V3ArgsE"),
    ("coupled pair / K v + M w (NBC=2)", "_ZN5spirk4k_v3ILi4ELi8ELi8ELi0ELi4ELi2EEEvNS_6V3ArgsE"),
    ("coupled pair, fused Chebyshev step (complex level operators, NBC=2)", "_ZN5spirk4k_v3ILi4ELi8ELi8ELi3ELi4ELi2EEEvNS_6V3ArgsE"),
    ("coupled pair, residual (NBC=2)", "_ZN5spirk4k_v3ILi4ELi8ELi8ELi1ELi4ELi2EEEvNS_6V3ArgsE"),
]
KEEP = ["DFMA", "DMUL", "DADD", "LDCU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "UTMALDG", "SYNCS", "BAR", "MEMBAR", "ATOMG", "CCTL", "DMMA", "HMMA"]


def sass(fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], capture_output=True, text=True).stdout
    return [l for l in out.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", l)]


def main():
    print("cuobjdump -sass of dealii_spirk_b200/libspirk_b200.so (final build of the round), per hot kernel: instruction count, opcode histogram of the Blackwell-relevant mnemonics")
    print("(UTMALDG = cp.async.bulk.tensor TMA loads, SYNCS = mbarrier ops, DFMA/DMUL/DADD = FP64 pipe; no DMMA/HMMA: the path is FP64 banded sweeps, HBM/FP64 co-limited;")
    print(" LDL/STL = register spills: loop-invariant values stored once per work item and reloaded per plane)\n")
    for title, fun in KERNELS:
        lines = sass(fun)
        ops = Counter()
        for l in lines:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
            if m:
                ops[m.group(1)] += 1
        print(f"== {title}\n   {fun}: {len(lines)} SASS instructions ({len(lines) * 16 // 1024} KB)")
        for k, v in sorted(((k, v) for k, v in ops.items() if k in KEEP), key=lambda kv: -kv[1]):
            print(f"   {v:6d} {k}")
    print("\nexcerpt: TMA issue + mbarrier arm / wait of the fused Chebyshev kernel (staged plane: 2 boxes, operands: 2 x 2 boxes)")
    n = 0
    for l in sass(KERNELS[0][1]):
        if "UTMALDG" in l or "SYNCS" in l:
            print(l.rstrip())
            n += 1
            if n >= 16:
                break


if __name__ == "__main__":
    sys.exit(main())
