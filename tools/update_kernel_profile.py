#!/usr/bin/env python
"""profiles/kernel_profile.json <- one `ncu --set full --import-source on` capture of the default W=5 perm kernel.

bench.py reports `roofline.traffic` and the ncu-verified executed IMAD.WIDE count only while the sha256 of the kernel
sources (bench.KERNEL_SOURCES) equals the one recorded here, so the numbers can never silently go stale.
usage: python tools/update_kernel_profile.py gpurun_out/prof_perm5.ncu-rep [profiles/rNN_ncu_....txt]"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def page(rep, name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr = rows[0]
    d = dict(zip(hdr, rows[2]))
    kernel = d.get("Kernel Name")
    grid = int(re.sub(r"[^0-9,]", "", d["Grid Size"]).split(",")[0])
    block = int(re.sub(r"[^0-9,]", "", d["Block Size"]).split(",")[0])
    unit = dict(zip(hdr, rows[1]))

    def nbytes(key):
        v, u = float(d[key]), unit[key].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

    dram = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
    states = grid * block  # one state per thread (the tail block may be partial: < 0.001 % at 2^26)
    src = list(csv.reader(io.StringIO(page(rep, "source"))))
    h = src[1]
    i_src, i_exec = h.index("Source"), h.index("Instructions Executed")
    wide = sum(float(r[i_exec]) for r in src[2:] if len(r) > i_exec and "IMAD.WIDE" in r[i_src])
    out = {
        "kernel": kernel, "grid": grid, "block": block, "states": states,
        "dram_bytes_per_launch": dram,
        "dram_bytes_per_launch_2p26": dram if states == (1 << 26) else None,
        "dram_bytes_per_state": dram / states,
        "executed_imad_wide_per_perm_ncu": wide / (states / 32),
        "fmaheavy_pct": float(d.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "nan")),
        "gpu_time": d.get("gpu__time_duration.sum", "nan") + " " + unit.get("gpu__time_duration.sum", ""),
        "ncu_file": sys.argv[2] if len(sys.argv) > 2 else os.path.basename(rep),
        "kernel_source_sha256": bench.kernel_source_hash(), "kernel_sources": bench.KERNEL_SOURCES,
    }
    path = os.path.join(ROOT, "profiles", "kernel_profile.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
