#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_attend.py tests/test_gpu_decode.py tests/test_step_golden.py tests/test_gpu_engine.py tests/test_gpu_score.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_handoff_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2_handoff_tests.log
run() {
  env $1 timeout 400 python bench.py --steps 64 --warmup 8 --no-whole-model --no-cpu-baseline --no-hit-sweep --no-reference-cuda --no-kernel-table > gpurun_out/r2_bench_ho_$2.json 2> gpurun_out/r2_bench_ho_$2.err; echo "bench $2 rc=$?"
}
run "SKV_SCORE_HANDOFF=1 SKV_OUT_FLAGS=1" both
run "SKV_SCORE_HANDOFF=0 SKV_OUT_FLAGS=1" outonly
run "SKV_SCORE_HANDOFF=1 SKV_OUT_FLAGS=0" scoreonly
run "SKV_SCORE_HANDOFF=0 SKV_OUT_FLAGS=0" none
python - <<'PY'
import json
for f in ("both", "outonly", "scoreonly", "none"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_ho_{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "hbm_first", round(d["hbm_first"]["value"], 1), round(d["hbm_first"]["us_per_layer_step"], 2))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2_bench_ho_{f}.err").read()[-1500:])
PY
