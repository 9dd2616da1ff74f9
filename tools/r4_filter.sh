#!/usr/bin/env bash
# Round-2 (session 4): log-domain candidate filter in the hinted selection -- tests + A/B of the bench (+ batch 8).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_decode.py tests/test_step_golden.py tests/test_gpu_fused_attend.py tests/test_gpu_engine.py tests/test_gpu_prefill.py tests/test_gpu_reference_python.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r4_filter_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r4_filter_tests.log
F="--steps 64 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model --no-cpu-baseline --no-reference-cuda"
for tag in on1 off1 on2 off2; do
  case $tag in on*) v=1;; *) v=0;; esac
  SKV_SELECT_FILTER=$v timeout 600 python bench.py $F > gpurun_out/r4_fl_$tag.json 2> gpurun_out/r4_fl_$tag.err; echo "$tag rc=$?"
done
for tag in b8on b8off; do
  case $tag in *on) v=1;; *) v=0;; esac
  SKV_SELECT_FILTER=$v timeout 600 python bench.py $F --batch 8 --layers 8 > gpurun_out/r4_fl_$tag.json 2> gpurun_out/r4_fl_$tag.err; echo "$tag rc=$?"
done
python - <<'PY'
import json
for tag in ("on1", "off1", "on2", "off2", "b8on", "b8off"):
    try:
        d = json.loads(open(f"gpurun_out/r4_fl_{tag}.json").read().strip().splitlines()[-1]); hb = d.get("hbm_first") or {}
        k = d["roofline"].get("kernels", {})
        print(tag, "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "hbm_first", hb.get("value") and round(hb["value"], 1), hb.get("us_per_layer_step") and round(hb["us_per_layer_step"], 2),
              {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k})
    except Exception as e:
        print(tag, "failed", e); print(open(f"gpurun_out/r4_fl_{tag}.err").read()[-1500:])
PY
SKV_LIB=shadowkv_b200/lib/libshadowkv_b200_trace.so timeout 300 python tools/trace_fused.py --hbm > gpurun_out/r4_trace_filter_hbm.txt 2>&1; cut -c1-260 gpurun_out/r4_trace_filter_hbm.txt | sed -n 5,6p
