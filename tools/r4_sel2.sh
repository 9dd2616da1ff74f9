#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_decode.py tests/test_step_golden.py tests/test_gpu_fused_attend.py tests/test_gpu_engine.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r4_sel2_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r4_sel2_tests.log
F="--steps 64 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model --no-cpu-baseline --no-reference-cuda"
for tag in a b; do
  timeout 600 python bench.py $F > gpurun_out/r4_sel2_$tag.json 2> gpurun_out/r4_sel2_$tag.err; echo "$tag rc=$?"
done
python - <<'PY'
import json
for tag in ("a", "b"):
    try:
        d = json.loads(open(f"gpurun_out/r4_sel2_{tag}.json").read().strip().splitlines()[-1]); hb = d.get("hbm_first") or {}
        k = d["roofline"].get("kernels", {})
        print(tag, "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "hbm_first", hb.get("value") and round(hb["value"], 1), hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2),
              {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k})
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r4_sel2_{tag}.err").read()[-1500:])
PY
SKV_LIB=shadowkv_b200/lib/libshadowkv_b200_trace.so timeout 300 python tools/trace_fused.py --hbm > gpurun_out/r4_trace_sel2_hbm.txt 2>&1; cut -c1-260 gpurun_out/r4_trace_sel2_hbm.txt | sed -n 2,7p
