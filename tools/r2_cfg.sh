#!/usr/bin/env bash
# One bench.py line on N GPUs of one box for a named BASELINE config: tools/r2_cfg.sh N name [bench.py args...]
N=$1; NAME=$2; shift 2
mkdir -p gpurun_out
EXTRA="--no-hit-sweep --no-head-split --no-whole-model --no-cpu-baseline --no-reference-cuda"
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 $EXTRA "$@" > gpurun_out/r2_$NAME.json 2> gpurun_out/r2_$NAME.err; echo "$NAME rc=$?"
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 20 --warmup 5 $EXTRA "$@" > gpurun_out/r2_$NAME.json 2> gpurun_out/r2_$NAME.err; echo "$NAME rc=$?"
fi
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_$NAME.json").read().strip().splitlines()[-1])
    k = d.get("kernels") or {}
    print("$NAME", "N", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "us/layer", round(d["us_per_layer_step"], 2),
          "e2e", round(d["e2e"]["value"], 1), "hit", d["chunk_cache"], "setup_s", d.get("setup_s"))
    print("  hbm_first", {x: (round(v, 2) if isinstance(v, float) else v) for x, v in (d.get("hbm_first") or {}).items() if x in ("value", "us_per_layer_step", "policy", "error")})
    print("  kernels", {x: {y: round(z, 3) for y, z in v.items() if isinstance(z, float)} for x, v in k.items() if x in ("score", "select", "attention")})
    print("  pcie_concurrent", d.get("pcie_concurrent"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_$NAME.err").read()[-2500:])
PY
