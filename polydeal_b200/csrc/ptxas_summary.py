#!/usr/bin/env python3
"""Summarise `nvcc -Xptxas -v` logs: kernel, registers, spills, static smem."""
import re
import subprocess
import sys

for path in sys.argv[1:]:
    lines = open(path).read().splitlines()
    cur = None
    for l in lines:
        m = re.search(r"Compiling entry function '([^']+)'", l)
        if m:
            cur = subprocess.run(["c++filt", m[1].split("__", 3)[-1] if False else m[1]], capture_output=True, text=True).stdout.strip()
            k = re.search(r"(k_\w+)(<[^>]*>)?", cur)
            cur = k[0] if k else cur[-60:]
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", l)
        if m and cur:
            stack = m.groups()
            continue
        m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes smem)?", l)
        if m and cur:
            print(f"{cur:55s} regs={m[1]:>3s} stack={stack[0]} spill_st={stack[1]} spill_ld={stack[2]}")
            cur = None
