#!/usr/bin/env python
"""bench.py -- Hades252 W=5 permutations per second on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2-states L]
  (N > 1: launched by torchrun, one rank per GPU; rank 0 prints ONE JSON line)

A "step" is one pass of the hot path (`perm_batch`) over one batch of synthetic states:
configs[1] of BASELINE.json = 2^26 width-5 states (10.7 GB) per GPU, generated on the device with
the splitmix64 generator of SURVEY.md 8(d).  `value` = whole-job perms/s with the states resident
in HBM; `e2e` = the same through the reference-facing call `hades_perm_batch` (C ABI) on pinned
HOST buffers, H2D and D2H inside the timed region.  Scaling is weak (fixed states per GPU, states
are independent, no data-path collective).

Roofline: the path is bound by the integer-multiply pipe, so
`roofline.bound = "int_mul"`, unit Tprod/s (1e12 32x32->64 limb products per second):
  achieved = perms/s x 268192 algorithmic limb-products per perm,
  peak     = hades_imad_peak microbenchmark measured live on the same device (best variant).
The HBM view (320 B/perm against MEASURED_PEAKS.json hbm_gbs) is reported beside it.

`--impl reference` times the CPU restatement of `ScalarStrategy::perm` (oracle/hades_cpu.c, the
"port": no Rust toolchain exists in this image) on all host threads, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH = 5
LIMB_PRODUCTS_PER_PERM = 268192   # 1972 Fr mul x 136 (8-limb CIOS), SURVEY.md 8(d)
HBM_BYTES_PER_PERM = 2 * 32 * WIDTH
SEED = 0x4861646573323532
METRIC = "hades252_w5_perms_per_sec"
UNIT = "perms/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-states", type=int, default=26, help="states per GPU per step (default 2^26)")
    ap.add_argument("--log2-e2e-states", type=int, default=None, help="states per GPU for the host e2e leg")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--variant", default=None, help="algo,regs kernel variant (default: library default)")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------ CPU legs
def cpu_rate(target_seconds: float, threads: int | None = None):
    """perms/s of the CPU port on `threads` host threads over a bounded sample of the workload."""
    import numpy as np
    from oracle import cpu_oracle
    threads = threads or cpu_oracle.host_threads()
    probe = 1 << 12
    s = cpu_oracle.gen_elems(0, WIDTH * probe, SEED).reshape(probe, WIDTH, 4)
    t = time.perf_counter(); cpu_oracle.perm_batch(s, WIDTH, threads); dt = time.perf_counter() - t
    n = int(min(1 << 20, max(1 << 12, probe / dt * target_seconds)))
    n = 1 << (n.bit_length() - 1)
    s = cpu_oracle.gen_elems(0, WIDTH * n, SEED).reshape(n, WIDTH, 4)
    best = 0.0
    for _ in range(2):
        t = time.perf_counter(); cpu_oracle.perm_batch(s, WIDTH, threads); dt = time.perf_counter() - t
        best = max(best, n / dt)
    return best, threads, n


def run_reference(args):
    """Reference arm: CPU restatement of ScalarStrategy::perm, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import cpu_oracle
    threads = cpu_oracle.host_threads()
    probe = 1 << 12
    s = cpu_oracle.gen_elems(0, WIDTH * probe, SEED).reshape(probe, WIDTH, 4)
    t = time.perf_counter(); cpu_oracle.perm_batch(s, WIDTH, threads); dt = time.perf_counter() - t
    total_steps = max(1, args.steps + args.warmup)
    n = int(min(1 << 20, max(1 << 12, probe / dt * (60.0 / total_steps))))
    n = 1 << (n.bit_length() - 1)
    s = cpu_oracle.gen_elems(0, WIDTH * n, SEED).reshape(n, WIDTH, 4)
    for _ in range(args.warmup):
        cpu_oracle.perm_batch(s, WIDTH, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle.perm_batch(s, WIDTH, threads)
    el = time.perf_counter() - t0
    value = n * args.steps / el
    sample = f"2^{n.bit_length() - 1} of the 2^{args.log2_states} synthetic states per step"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (255-bit modular integer)",
        "data": "synthetic",
        "config": {"workload": f"batched perm, width 5, 2^{args.log2_states} states per GPU (BASELINE configs[1])",
                   "sample": sample, "seed": hex(SEED)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "C restatement of ScalarStrategy::perm (oracle/hades_cpu.c, pthreads); "
                                 "no Rust toolchain in this image, reference not runnable"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(pw), samples=len(sm),
                       reasons=sorted(reasons))
        return out


# ------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from hades252_b200 import CudaStrategy
    strat = CudaStrategy([local])
    if args.variant:
        algo, regs = (int(x) for x in args.variant.split(","))
        strat.set_variant(algo, regs)
    n = 1 << args.log2_states
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # roofline denominator, measured live on this device
    peaks = {v: strat.imad_peak(v) for v in range(4)}
    names = {0: "imad_wide_x_carry_chain", 1: "imad_wide_carry_out_only", 2: "imad_lo32_half_product_context_only",
             3: "imad_lo_plus_imad_hi_pair"}
    p_mul32 = max(peaks[0], peaks[1], peaks[3])  # forms that deliver a full 64-bit multiply-accumulate
    info = strat.kernel_info("perm")

    # ---- device-resident leg: states generated on device, permuted in place K times -------------
    states = torch.empty(n * WIDTH * 4, dtype=torch.int64, device="cuda")
    strat.gen_elems_device(states.data_ptr(), rank * n * WIDTH, n * WIDTH, SEED, sptr)
    for _ in range(args.warmup):
        strat.perm_batch_device(states.data_ptr(), n, sptr)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = strat.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record(stream)
    for k in range(args.steps):
        strat.perm_batch_device(states.data_ptr(), n, sptr)
        ev[k + 1].record(stream)
    barrier()
    launches = strat.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = torch.tensor([ev[0].elapsed_time(ev[-1])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * n * args.steps / (total_ms * 1e-3)
    kernel_ms = statistics.mean(step_ms)

    # correctness spot check (outside the timed region): digest of the final states + a sample
    dig = torch.zeros(4, dtype=torch.int64, device="cuda")
    strat.digest_device(states.data_ptr(), 0, n * WIDTH * 4, dig.data_ptr(), sptr)
    torch.cuda.synchronize()
    digest = [hex(int(x)) for x in dig.cpu().numpy().view(np.uint64)]
    del states
    torch.cuda.empty_cache()

    # ---- end-to-end leg: pinned host buffers through the reference-facing C-ABI call --------------
    e2e = None
    if not args.no_e2e:
        l2e = args.log2_e2e_states if args.log2_e2e_states is not None else args.log2_states
        host = None
        while l2e >= 16:
            try:
                host = torch.empty((1 << l2e) * WIDTH * 4, dtype=torch.int64, pin_memory=True)
                break
            except RuntimeError:
                l2e -= 1
        ne = 1 << l2e
        # fill the host buffer with synthetic states (generated on device, copied once, untimed)
        tmp = torch.empty(min(ne, 1 << 22) * WIDTH * 4, dtype=torch.int64, device="cuda")
        chunk = tmp.numel() // (WIDTH * 4)
        for off in range(0, ne, chunk):
            strat.gen_elems_device(tmp.data_ptr(), (rank * ne + off) * WIDTH, chunk * WIDTH, SEED, sptr)
            torch.cuda.synchronize()
            host[off * WIDTH * 4:(off + chunk) * WIDTH * 4].copy_(tmp)
        del tmp
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(min(args.warmup, 1) or 1):
            strat.perm_batch_ptr(host.data_ptr(), ne)
        barrier()
        l0 = strat.launch_count
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            strat.perm_batch_ptr(host.data_ptr(), ne)   # synchronous: returns with outputs in host memory
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e_launches = strat.launch_count - l0
        e2e = {"value": world * ne * e2e_steps / float(el.item()), "unit": UNIT,
               "h2d_bytes_per_step": ne * HBM_BYTES_PER_PERM // 2 * world, "d2h_bytes_per_step": ne * HBM_BYTES_PER_PERM // 2 * world,
               "states_per_gpu_per_step": ne, "steps": e2e_steps, "host_memory": "pinned",
               "api": "hades_perm_batch (C ABI, chunked H2D/kernel/D2H pipeline)", "gpu_launches": e2e_launches}
        del host

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        rate, cores, sample_n = cpu_rate(args.cpu_seconds)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"2^{sample_n.bit_length() - 1} of the synthetic states, best of 2",
                        "note": "C restatement of ScalarStrategy::perm (oracle/hades_cpu.c, pthreads over host cores)"}

    if rank == 0:
        pk, pk_src = measured_peaks()
        per_gpu = value / world
        achieved = per_gpu * LIMB_PRODUCTS_PER_PERM / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (255-bit modular integer)", "data": "synthetic",
            "config": {"workload": f"batched perm, width 5, 2^{args.log2_states} states per GPU (BASELINE configs[1])",
                       "states_per_gpu": n, "bytes_per_gpu": n * WIDTH * 32, "seed": hex(SEED), "in_place": True,
                       "l2_policy": "inputs (10.7 GB) larger than L2", "parallelism": f"dp{world} (independent states, no collective)"},
            "roofline": {"bound": "int_mul", "achieved": achieved, "peak": p_mul32 / 1e12, "unit": "Tprod/s",
                         "frac": achieved / (p_mul32 / 1e12), "traffic": None,
                         "kernel": "perm_batch_kernel (width 5)", "variant": args.variant or "default", "kernel_ms": kernel_ms,
                         "algorithmic_products_per_perm": LIMB_PRODUCTS_PER_PERM,
                         "peak_source": "hades_imad_peak live on this device",
                         "peak_variants_Tprod_s": {names[v]: peaks[v] / 1e12 for v in peaks}},
            "roofline_hbm": {"achieved_gbs": per_gpu * HBM_BYTES_PER_PERM / 1e9, "peak_gbs": pk.get("hbm_gbs"),
                             "peak_source": pk_src, "frac": per_gpu * HBM_BYTES_PER_PERM / 1e9 / pk.get("hbm_gbs", 1)},
            "kernel_info": info, "gpu_launches": launches, "clocks": clocks, "digest": digest,
            "e2e": e2e, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(out))
    strat.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
