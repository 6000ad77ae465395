#!/usr/bin/env python
"""Summarise hades252_b200/lib/build.log (nvcc -Xptxas -v): registers / stack / spills per kernel.
usage: python tools/build_report.py [substring ...]   -- only kernels whose mangled name contains every substring"""
import os
import re
import sys

LOG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hades252_b200", "lib", "build.log")


def main():
    want = sys.argv[1:]
    name = None
    stack = spill_st = spill_ld = 0
    for line in open(LOG):
        m = re.search(r"Function properties for (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            stack, spill_st, spill_ld = map(int, m.groups())
            continue
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            short = re.sub(r"_ZN5hades\d+_GLOBAL__N__[0-9a-f]+_\d+_", "", name)
            if all(w in short for w in want):
                print(f"{short:90s} regs {int(m.group(1)):3d}  stack {stack:4d}  spill st/ld {spill_st}/{spill_ld}")
            name = None


if __name__ == "__main__":
    main()
