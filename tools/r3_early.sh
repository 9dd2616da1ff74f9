#!/usr/bin/env bash
# A/B: split CTAs of the fused launch request their first K / V tiles ahead of the row hand-over (SKV_FA_EARLY)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_fused_attend.py tests/test_step_golden.py tests/test_gpu_decode.py tests/test_gpu_engine.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r3_early_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r3_early_tests.log
run() {  # name batch layers env...
  local name=$1 b=$2 l=$3; shift 3
  env "$@" timeout 600 python bench.py --batch $b --layers $l --steps 64 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model \
     --no-cpu-baseline --no-reference-cuda --no-kernel-table > gpurun_out/r3_$name.json 2> gpurun_out/r3_$name.err; echo "$name rc=$?"
}
run early1_b1 1 32 SKV_FA_EARLY=1
run early0_b1 1 32 SKV_FA_EARLY=0
run early1_b1b 1 32 SKV_FA_EARLY=1
run early0_b1b 1 32 SKV_FA_EARLY=0
run early1_b8 8 8 SKV_FA_EARLY=1
run early0_b8 8 8 SKV_FA_EARLY=0
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/r3_early*.json"), key=os.path.getmtime):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); hb = d.get("hbm_first") or {}
        print(os.path.basename(f), "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "hbm_first", hb.get("value") and round(hb["value"], 1),
              hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2))
    except Exception as e:
        print(f, "failed", e); print(open(f[:-4] + "err").read()[-800:])
PY
