#!/bin/bash
# GPU-box helper: parity suite + a short bench, prints the per-kernel times.  usage: tools/gpu_check.sh [tag] [bench args...]
TAG=${1:-chk}; shift || true
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 40 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("$TAG", "frames/s", round(d["value"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), "integrate us", round(d["roofline"]["us_per_launch"], 2), "frac",
      round(d["roofline"]["frac"], 4), {k: round(v["ms_per_launch"] * 1e3, 1) for k, v in d["kernels"].items()})
PY
