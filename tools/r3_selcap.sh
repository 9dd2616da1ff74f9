#!/usr/bin/env bash
# A/B: select_kernel register cap and shared-memory padding (CTAs per SM) at batch 1 / 2 / 4 / 8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_score.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -2
run() {  # name batch layers env...
  local name=$1 b=$2 l=$3; shift 3
  env "$@" timeout 600 python bench.py --batch $b --layers $l --steps 48 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model \
     --no-cpu-baseline --no-reference-cuda > gpurun_out/r3_$name.json 2> gpurun_out/r3_$name.err; echo "$name rc=$?"
}
run b1_cap1 1 32 SKV_SELECT_CAP=1
run b1_cap0 1 32 SKV_SELECT_CAP=0
run b1_cap0_pad64 1 32 SKV_SELECT_CAP=0 SKV_SELECT_PAD_KB=64
run b1_cap0_pad100 1 32 SKV_SELECT_CAP=0 SKV_SELECT_PAD_KB=100
run b2_cap1 2 16 SKV_SELECT_CAP=1
run b2_cap0 2 16 SKV_SELECT_CAP=0
run b8_cap1 8 8 SKV_SELECT_CAP=1
run b8_cap0 8 8 SKV_SELECT_CAP=0
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/r3_b*.json"), key=os.path.getmtime):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); k = d["kernels"]; hb = d.get("hbm_first") or {}
        print(os.path.basename(f), "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "hbm_first", hb.get("value") and round(hb["value"], 1),
              hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2), {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k})
    except Exception as e:
        print(f, "failed", e); print(open(f[:-4] + "err").read()[-800:])
PY
