#!/usr/bin/env bash
# A/B of the single-launch V fetch + K rebuild (default) against the two launches on two streams it replaces
# (SKV_FUSED_FETCH=0), at batch 1 (32 layers) and batch 8 (8 layers); about 2.5 GPU-minutes:
#   gpurun --timeout 1200 -- 'bash tools/gpu_fused_check.sh'
mkdir -p gpurun_out
T="timeout 400 python -m pytest -q -m gpu -x --no-header -p no:cacheprovider"
$T tests/test_gpu_decode.py tests/test_gpu_engine.py > gpurun_out/fused_tests.log 2>&1; echo "fused tests rc=$?"; tail -3 gpurun_out/fused_tests.log
SKV_FUSED_FETCH=0 $T tests/test_gpu_decode.py tests/test_gpu_engine.py > gpurun_out/twostream_tests.log 2>&1; echo "two-stream tests rc=$?"; tail -3 gpurun_out/twostream_tests.log
B="timeout 600 python bench.py --no-whole-model --no-cpu-baseline --hbm-store-ways 0"
$B > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "bench fused rc=$?"
SKV_ATTN_PDL=0 $B --no-kernel-table > gpurun_out/bench_fused_nopdl.json 2> gpurun_out/bench_fused_nopdl.err; echo "bench fused, attention not programmatic rc=$?"
SKV_FUSED_FETCH=0 $B --no-kernel-table > gpurun_out/bench_twostream.json 2> gpurun_out/bench_twostream.err; echo "bench two-stream rc=$?"
$B --no-kernel-table > gpurun_out/bench_fused2.json 2> gpurun_out/bench_fused2.err; echo "bench fused (repeat) rc=$?"
B8="$B --batch 8 --layers 8 --steps 48 --no-kernel-table"
$B8 > gpurun_out/bench_b8_fused.json 2> gpurun_out/bench_b8_fused.err; echo "batch 8 fused rc=$?"
SKV_FUSED_FETCH=0 $B8 > gpurun_out/bench_b8_twostream.json 2> gpurun_out/bench_b8_twostream.err; echo "batch 8 two-stream rc=$?"
python - <<'PY'
import json
for n in ("fused", "fused_nopdl", "twostream", "fused2", "b8_fused", "b8_twostream"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "us/layer", round(d["us_per_layer_step"], 2),
              "launches", d["gpu_launches"], d["config"]["fetch_rebuild_launch"])
        for h in ("h60", "h0"):
            if (d.get("kernels") or {}).get(h):
                print("  ", h, {k: round(v, 2) for k, v in d["kernels"][h].items()})
    except Exception as e:
        print(n, "failed", e)
PY
