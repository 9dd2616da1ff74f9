#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_score.py tests/test_gpu_decode.py tests/test_step_golden.py tests/test_gpu_fused_attend.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r3_check_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r3_check_tests.log
timeout 600 python bench.py --steps 64 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model --no-cpu-baseline --no-reference-cuda > gpurun_out/r3_spread.json 2> gpurun_out/r3_spread.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r3_spread.json").read().strip().splitlines()[-1]); k = d["kernels"]; hb = d.get("hbm_first") or {}
print("value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "hbm_first", hb.get("value") and round(hb["value"], 1),
      hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2), {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k})
PY
bash tools/r2_trace.sh
