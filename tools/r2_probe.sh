#!/usr/bin/env bash
# Round-2 probe: phase timelines of the selection / scoring / attention kernels (trace build) and the isolated
# scoring kernel at several grid sizes (what a fused score+select cluster kernel would have to live with).
mkdir -p gpurun_out
SKV_LIB=shadowkv_b200/lib/libshadowkv_b200_trace.so timeout 300 python tools/trace_kernels.py > gpurun_out/r2_trace.txt 2>&1; echo "trace rc=$?"
for c in 64 128 148 256 296; do
  SKV_SCORE_CTAS=$c timeout 300 python bench.py --steps 8 --warmup 3 --no-kernel-table --no-cpu-baseline --hbm-store-ways 0 --no-whole-model \
    > gpurun_out/r2_score_ctas_$c.json 2> gpurun_out/r2_score_ctas_$c.err; echo "ctas $c rc=$?"
done
python - <<'PY'
import json
for c in (64, 128, 148, 256, 296):
    try:
        d = json.loads(open(f"gpurun_out/r2_score_ctas_{c}.json").read().strip().splitlines()[-1])
        print(c, "score us", round(d["kernels"]["score"]["us"], 2), "value", round(d["value"], 1))
    except Exception as e:
        print(c, "failed", e)
PY
tail -40 gpurun_out/r2_trace.txt
