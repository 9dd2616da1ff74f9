#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_score.py -x -q -m gpu --no-header -p no:cacheprovider -k "candidate_filter or fast_path" > gpurun_out/r4_filter2_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r4_filter2_tests.log
# the 512K configuration on one GPU (4 layers): the selection with 1024-thread CTAs and the filter
F="--steps 20 --warmup 5 --no-hit-sweep --no-head-split --no-whole-model --no-cpu-baseline --no-reference-cuda --model llama-3-8b-1m --ctx 524288 --batch 1 --layers 4"
for v in 1 0; do
  SKV_SELECT_FILTER=$v timeout 600 python bench.py $F > gpurun_out/r4_fl512_$v.json 2> gpurun_out/r4_fl512_$v.err; echo "512K filter=$v rc=$?"
done
python - <<'PY'
import json
for tag in ("1", "0"):
    try:
        d = json.loads(open(f"gpurun_out/r4_fl512_{tag}.json").read().strip().splitlines()[-1]); hb = d.get("hbm_first") or {}
        k = d["roofline"].get("kernels", {})
        print("filter", tag, "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2),
              "hbm_first", hb.get("value") and round(hb["value"], 1), hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2),
              {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k})
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r4_fl512_{tag}.err").read()[-1500:])
PY
