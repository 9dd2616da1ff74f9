#!/usr/bin/env bash
# Row hand-over between the selection and the fused fetch launch: parity tests, then the bench A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_attend.py tests/test_gpu_decode.py tests/test_gpu_engine.py tests/test_step_golden.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_rowflags_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2_rowflags_tests.log
for f in 1 0; do
  SKV_ROW_FLAGS=$f timeout 400 python bench.py --steps 64 --warmup 8 --no-whole-model --no-cpu-baseline --no-hit-sweep --no-reference-cuda \
    > gpurun_out/r2_bench_rf$f.json 2> gpurun_out/r2_bench_rf$f.err; echo "bench row_flags=$f rc=$?"
done
python - <<'PY'
import json
for f in (1, 0):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_rf{f}.json").read().strip().splitlines()[-1])
        print("row_flags", f, "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "hbm_first", round(d["hbm_first"]["value"], 1), round(d["hbm_first"]["us_per_layer_step"], 2))
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2_bench_rf{f}.err").read()[-1500:])
PY
