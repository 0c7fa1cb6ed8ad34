#!/usr/bin/env python
"""Dynamic instruction mix of the first kernel in an .ncu-rep (read here, without a GPU): warp instructions executed
per SASS opcode, and per basic region between branch targets if asked.

usage: python tools/sass_mix.py gpurun_out/prof.ncu-rep [rays_per_launch] [--dump]
"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    # first kernel only
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
    rd = csv.DictReader(io.StringIO("\n".join(lines[start:end])))
    return list(rd)


def main():
    rep = sys.argv[1]
    rays = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
    rs = rows(rep)
    mix, tot = Counter(), 0
    for r in rs:
        src = r["Source"].strip()
        m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
        op = m.group(2) if m else "?"
        n = int(r["Instructions Executed"])
        mix[op] += n
        tot += n
    print(f"total warp instructions {tot:,}" + (f" = {tot / rays:.2f} per ray ({tot * 32 / rays:.0f} thread-slots)" if rays else ""))
    for op, n in mix.most_common(40):
        print(f"{op:12s} {n:14,d} {100 * n / tot:5.1f}%" + (f"  {n * 32 / rays:6.1f}/ray" if rays else ""))
    if "--dump" in sys.argv:
        for r in rs:
            print(f'{int(r["Instructions Executed"]):12d} {int(r["# Samples"]):6d}  {r["Source"].strip()}')


if __name__ == "__main__":
    main()
