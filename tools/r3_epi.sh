#!/usr/bin/env bash
# tests + A/B: epilogue warpgroups of the scoring kernel (SKV_SCORE_EPI=2/4) and L2 prefetch of the tiles beyond its ring (SKV_SCORE_L2PF)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_dense.py tests/test_step_golden.py tests/test_gpu_decode.py tests/test_gpu_engine.py tests/test_gpu_prefill.py tests/test_ref_golden.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r3_epi_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r3_epi_tests.log
SKV_SCORE_EPI=4 timeout 900 python -m pytest tests/test_gpu_score.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -2
run() {  # name batch layers env...
  local name=$1 b=$2 l=$3; shift 3
  env "$@" timeout 600 python bench.py --batch $b --layers $l --steps 64 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model \
     --no-cpu-baseline --no-reference-cuda > gpurun_out/r3_$name.json 2> gpurun_out/r3_$name.err; echo "$name rc=$?"
}
run epi4_pf1 1 32 SKV_SCORE_EPI=4 SKV_SCORE_L2PF=1
run epi2_pf0 1 32 SKV_SCORE_EPI=2 SKV_SCORE_L2PF=0
run epi4_pf0 1 32 SKV_SCORE_EPI=4 SKV_SCORE_L2PF=0
run epi2_pf1 1 32 SKV_SCORE_EPI=2 SKV_SCORE_L2PF=1
run epi4_pf1b 1 32 SKV_SCORE_EPI=4 SKV_SCORE_L2PF=1
run epi2_pf0b 1 32 SKV_SCORE_EPI=2 SKV_SCORE_L2PF=0
run epi4_b8 8 8 SKV_SCORE_EPI=4
run epi2_b8 8 8 SKV_SCORE_EPI=2
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/r3_epi*.json"), key=os.path.getmtime):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); k = d["kernels"]; hb = d.get("hbm_first") or {}
        print(os.path.basename(f), "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "hbm_first", hb.get("value") and round(hb["value"], 1),
              hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2), {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k})
    except Exception as e:
        print(f, "failed", e); print(open(f[:-4] + "err").read()[-800:])
PY
SKV_LIB=shadowkv_b200/lib/libshadowkv_b200_trace.so timeout 300 python tools/trace_fused.py --hbm 2>&1 | cut -c1-200 | head -6
