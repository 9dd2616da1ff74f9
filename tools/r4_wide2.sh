#!/usr/bin/env bash
# 200K context, G = 8 (Yi shape): CTA width of the selection with the filter
mkdir -p gpurun_out
F="--steps 20 --warmup 5 --no-hit-sweep --no-head-split --no-whole-model --no-cpu-baseline --no-reference-cuda --model yi-9b-200k --ctx 204800 --batch 1 --layers 4"
run() { tag=$1; shift; env "$@" timeout 600 python bench.py $F > gpurun_out/r4_wide2_$tag.json 2> gpurun_out/r4_wide2_$tag.err; echo "$tag rc=$?"; }
run t512 SKV_X=0
run t256 SKV_SELECT_WIDE_FROM=1000000
F="--steps 20 --warmup 5 --no-hit-sweep --no-head-split --no-whole-model --no-cpu-baseline --no-reference-cuda --model llama-3-8b-1m --ctx 524288 --batch 1 --layers 4"
run l512k SKV_X=0
python - <<'PY'
import json
for tag in ("t512", "t256", "l512k"):
    try:
        d = json.loads(open(f"gpurun_out/r4_wide2_{tag}.json").read().strip().splitlines()[-1]); hb = d.get("hbm_first") or {}
        k = d["roofline"].get("kernels", {})
        print(tag, "us/layer", round(d["us_per_layer_step"], 2), "hbm_first", hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2),
              {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k})
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r4_wide2_{tag}.err").read()[-1500:])
PY
