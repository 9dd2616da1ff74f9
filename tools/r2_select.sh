#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_score.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_select_tests.log 2>&1; echo "score/select tests rc=$?"
tail -15 gpurun_out/r2_select_tests.log
timeout 900 python -m pytest tests/test_step_golden.py tests/test_gpu_decode.py tests/test_gpu_engine.py tests/test_gpu_fused_attend.py tests/test_ref_golden.py tests/test_gpu_prefill.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_select_regress.log 2>&1; echo "regress rc=$?"
tail -4 gpurun_out/r2_select_regress.log
for f in 1 0; do
  SKV_FUSED_SELECT=$f timeout 600 python bench.py --steps 64 --warmup 8 --no-whole-model --no-cpu-baseline --no-hit-sweep \
    > gpurun_out/r2_bench_sel$f.json 2> gpurun_out/r2_bench_sel$f.err; echo "bench fused_select=$f rc=$?"
done
python - <<'PY'
import json
for f in (1, 0):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_sel{f}.json").read().strip().splitlines()[-1])
        k = d["kernels"]
        print("fused_select", f, "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "hbm_first", round(d["hbm_first"]["value"], 1), round(d["hbm_first"]["us_per_layer_step"], 2),
              {x: round(k[x]["us"], 2) for x in ("score", "select", "score_select") if x in k})
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2_bench_sel{f}.err").read()[-1500:])
PY
