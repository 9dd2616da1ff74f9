#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_attend.py tests/test_gpu_decode.py tests/test_step_golden.py tests/test_gpu_engine.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_qk_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r2_qk_tests.log
SKV_LIB=shadowkv_b200/lib/libshadowkv_b200_trace.so timeout 300 python tools/trace_fused.py --hbm > gpurun_out/r2_trace_fused_hbm.txt 2>&1; echo "trace hbm rc=$?"
cut -c1-190 gpurun_out/r2_trace_fused_hbm.txt | head -12
timeout 400 python bench.py --steps 64 --warmup 8 --no-whole-model --no-cpu-baseline --no-hit-sweep --no-reference-cuda > gpurun_out/r2_bench_qk.json 2> gpurun_out/r2_bench_qk.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_qk.json").read().strip().splitlines()[-1])
k = d["kernels"]
print("value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "e2e", round(d["e2e"]["value"], 1),
      "hbm_first", round(d["hbm_first"]["value"], 1), round(d["hbm_first"]["us_per_layer_step"], 2),
      {x: {kk: round(vv, 2) for kk, vv in k[x].items() if kk.endswith("_us")} for x in ("h0", "h60", "h100") if x in k})
PY
