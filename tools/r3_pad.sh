#!/usr/bin/env bash
# A/B: shared-memory padding of select_kernel (CTAs of it per SM / co-residency with its neighbours) at batch 1; dense batch-64 tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -3
run() {  # name batch layers env...
  local name=$1 b=$2 l=$3; shift 3
  env "$@" timeout 600 python bench.py --batch $b --layers $l --steps 64 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model \
     --no-cpu-baseline --no-reference-cuda > gpurun_out/r3_$name.json 2> gpurun_out/r3_$name.err; echo "$name rc=$?"
}
run p0 1 32 SKV_SELECT_PAD_KB=0
run p76 1 32 SKV_SELECT_PAD_KB=76
run p84 1 32 SKV_SELECT_PAD_KB=84
run p92 1 32 SKV_SELECT_PAD_KB=92
run p100 1 32 SKV_SELECT_PAD_KB=100
run p0b 1 32 SKV_SELECT_PAD_KB=0
run b2_p0 2 16 SKV_SELECT_PAD_KB=0
run b2_p84 2 16 SKV_SELECT_PAD_KB=84
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/r3_p*.json") + glob.glob("gpurun_out/r3_b2_p*.json"), key=os.path.getmtime):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); k = d["kernels"]; hb = d.get("hbm_first") or {}
        print(os.path.basename(f), "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "hbm_first", hb.get("value") and round(hb["value"], 1),
              hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2), {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k})
    except Exception as e:
        print(f, "failed", e); print(open(f[:-4] + "err").read()[-800:])
PY
