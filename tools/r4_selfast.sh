#!/usr/bin/env bash
# Round-2 (session 4): single-CTA hinted selection in front of the cluster kernel -- tests + A/B of the bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_score.py -x -q -m gpu -s --no-header -p no:cacheprovider > gpurun_out/r4_selfast_score_tests.log 2>&1; echo "score tests rc=$?"; grep -a "fast-path status" gpurun_out/r4_selfast_score_tests.log | cut -c1-400; tail -4 gpurun_out/r4_selfast_score_tests.log
timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_step_golden.py tests/test_gpu_fused_attend.py tests/test_gpu_engine.py tests/test_gpu_prefill.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r4_selfast_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r4_selfast_tests.log
F="--steps 64 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model --no-cpu-baseline --no-reference-cuda"
for tag in on1 off1 on2 off2; do
  case $tag in on*) v=1;; *) v=0;; esac
  SKV_SELECT_FAST=$v timeout 600 python bench.py $F > gpurun_out/r4_sf_$tag.json 2> gpurun_out/r4_sf_$tag.err; echo "$tag rc=$?"
done
python - <<'PY'
import json
for tag in ("on1", "off1", "on2", "off2"):
    try:
        d = json.loads(open(f"gpurun_out/r4_sf_{tag}.json").read().strip().splitlines()[-1]); hb = d.get("hbm_first") or {}
        k = d["roofline"].get("kernels", {})
        print(tag, "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "hbm_first", hb.get("value") and round(hb["value"], 1), hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2),
              {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k}, "launches", d.get("gpu_launches"))
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r4_sf_{tag}.err").read()[-1500:])
PY
