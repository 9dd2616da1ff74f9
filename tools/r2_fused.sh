#!/usr/bin/env bash
# Round-2: the fused fetch + rebuild + attention launch -- tests, a compute-sanitizer pass, and the bench A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_attend.py -x -q -s -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_fused_tests.log 2>&1; echo "fused tests rc=$?"
grep -v "^$" gpurun_out/r2_fused_tests.log | tail -25
if [ "$1" = "san" ]; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_fused_attend.py::test_fused_launch_matches_oracle" -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_fused_memcheck.log 2>&1; echo "memcheck rc=$?"
  tail -5 gpurun_out/r2_fused_memcheck.log
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest "tests/test_gpu_fused_attend.py::test_fused_launch_matches_oracle" -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_fused_racecheck.log 2>&1; echo "racecheck rc=$?"
  tail -5 gpurun_out/r2_fused_racecheck.log
fi
timeout 600 python -m pytest tests/test_step_golden.py tests/test_gpu_decode.py tests/test_gpu_engine.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_fused_regress.log 2>&1; echo "regress rc=$?"
tail -5 gpurun_out/r2_fused_regress.log
for f in 1 0; do
  SKV_FUSED_ATTEND=$f timeout 600 python bench.py --steps 64 --warmup 8 --no-whole-model --no-cpu-baseline --no-hbm-first --no-hit-sweep \
    > gpurun_out/r2_bench_fused$f.json 2> gpurun_out/r2_bench_fused$f.err; echo "bench fused=$f rc=$?"
done
python - <<'PY'
import json
for f in (1, 0):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_fused{f}.json").read().strip().splitlines()[-1])
        k = d["kernels"]
        print("fused", f, "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "e2e", round(d["e2e"]["value"], 1),
              "hit", d["chunk_cache"]["hit_rate"], {x: {kk: round(vv, 2) for kk, vv in k[x].items() if kk.endswith("_us")} for x in ("h0", "h60", "h100") if x in k},
              "clocks", d["clocks"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/r2_bench_fused{f}.err").read()[-1500:])
PY
