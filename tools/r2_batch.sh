#!/usr/bin/env bash
mkdir -p gpurun_out
B=${1:-8}
timeout 900 python bench.py --batch $B --layers 8 --v-residency hbm --steps 32 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model --no-cpu-baseline --no-reference-cuda --no-hbm-first \
   > gpurun_out/r2_batch$B.json 2> gpurun_out/r2_batch$B.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_batch$B.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "us/layer", round(d["us_per_layer_step"],2), d["chunk_cache"])
for k, v in d["kernels"].items():
    print(" ", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a != "note" and a != "bound"})
PY
