#!/usr/bin/env python3
"""Registers / spills per kernel from a `nvcc -Xptxas -v` log:  python tools/ptxas_report.py LOG [name-filter]"""
import re
import subprocess
import sys

log = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for b in log.split("Compiling entry function '")[1:]:
    name = b.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
    if flt not in dem:
        continue
    regs = re.search(r"Used (\d+) registers", b)
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    sm = re.search(r"(\d+) bytes smem", b)
    print(f"{dem:110s} regs={regs.group(1) if regs else '?':>3} stack/spill={'/'.join(sp.groups()) if sp else '?'} "
          f"smem={sm.group(1) if sm else 0}")
