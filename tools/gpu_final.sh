#!/bin/bash
# What the driver does at round end, in one call.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; cut -c1-220 gpurun_out/bench_ref.json
python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print({k: d[k] for k in ("metric", "value", "ms_per_step", "n_gpus", "gpu_launches", "oracle_sample_match", "oracle_sample_states")})
print("e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "roofline", {k: d["roofline"][k] for k in ("bound", "achieved", "peak", "frac", "frac_pipe", "frac_algorithmic", "traffic")}, d["clocks"])
PY
