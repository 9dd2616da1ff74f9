#!/usr/bin/env bash
# bench.py under torchrun on N GPUs of one box, the way the driver launches it.
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "e2e", round(d["e2e"]["value"], 1))
    print("hbm_first", d["hbm_first"])
    print("pcie_concurrent", d["pcie_concurrent"])
    print("head_split", d["head_split"])
    print("by_hit", d["e2e"]["by_hit_rate"]["rows"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_bench_n$N.err").read()[-3000:])
PY
