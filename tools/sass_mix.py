#!/usr/bin/env python
"""Opcode mix of one kernel's SASS: whole kernel and its largest inner loop (the partial-round loop).
usage: python tools/sass_mix.py <object.o> <mangled-kernel-name-substring>"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    names = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
    funs = sorted(set(re.findall(r"\.text\.(\S+)", names)))
    fun = [f for f in funs if pat in f]
    if not fun:
        sys.exit(f"no kernel matching {pat}")
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun[0], obj], capture_output=True, text=True).stdout
    ins = []
    for line in sass.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())))
    best = None
    for a, t in ins:
        m = re.match(r"BRA(?:\.U)? (?:U?P\d, )?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and a - tgt < 0x8000 and (best is None or a - tgt > best[1] - best[0]):
                best = (tgt, a)
    print(fun[0])
    c = collections.Counter(t.split()[0] for a, t in ins)
    print(f"whole kernel: {len(ins)} instructions, {16 * len(ins)} bytes:", c.most_common(12))
    if best:
        c = collections.Counter(t.split()[0] for a, t in ins if best[0] <= a <= best[1])
        print(f"largest loop {best[0]:#x}..{best[1]:#x} = {best[1] - best[0]} bytes, {sum(c.values())} instructions:", c.most_common(14))


if __name__ == "__main__":
    main()
