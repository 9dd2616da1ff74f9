#!/usr/bin/env bash
# tests + A/B: threshold hint of the selection (SKV_SELECT_HINT)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_dense.py tests/test_gpu_fused_attend.py tests/test_step_golden.py tests/test_gpu_decode.py tests/test_gpu_engine.py tests/test_gpu_prefill.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r3_hint_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r3_hint_tests.log
run() {  # name batch layers env...
  local name=$1 b=$2 l=$3; shift 3
  env "$@" timeout 600 python bench.py --batch $b --layers $l --steps 64 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model \
     --no-cpu-baseline --no-reference-cuda > gpurun_out/r3_$name.json 2> gpurun_out/r3_$name.err; echo "$name rc=$?"
}
run hint1_b1 1 32 SKV_SELECT_HINT=1
run hint0_b1 1 32 SKV_SELECT_HINT=0
run hint1_b1b 1 32 SKV_SELECT_HINT=1
run hint0_b1b 1 32 SKV_SELECT_HINT=0
run hint1_b8 8 8 SKV_SELECT_HINT=1
run hint0_b8 8 8 SKV_SELECT_HINT=0
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/r3_hint*.json"), key=os.path.getmtime):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); k = d["kernels"]; hb = d.get("hbm_first") or {}
        print(os.path.basename(f), "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "hbm_first", hb.get("value") and round(hb["value"], 1),
              hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2), {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k})
    except Exception as e:
        print(f, "failed", e); print(open(f[:-4] + "err").read()[-800:])
PY
