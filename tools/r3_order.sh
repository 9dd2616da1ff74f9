#!/usr/bin/env bash
# A/B: CTA order of the fused fetch + rebuild + attention launch (SKV_FA_ORDER) at batch 8 / 4 / 1, V in HBM and on the host
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_score.py tests/test_gpu_fused_attend.py tests/test_step_golden.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -2
SKV_FA_ORDER=2 timeout 900 python -m pytest tests/test_gpu_fused_attend.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -2
run() {  # name batch layers residency env...
  local name=$1 b=$2 l=$3 r=$4; shift 4
  env "$@" timeout 600 python bench.py --batch $b --layers $l --v-residency $r --steps 32 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model \
     --no-cpu-baseline --no-reference-cuda --no-hbm-first --no-kernel-table > gpurun_out/r3_$name.json 2> gpurun_out/r3_$name.err; echo "$name rc=$?"
}
for o in 0 1 2; do
  run ord${o}_b8_hbm 8 8 hbm SKV_FA_ORDER=$o
  [ $o != 1 ] && run ord${o}_b8_host 8 8 host SKV_FA_ORDER=$o
  run ord${o}_b1_hbm 1 32 hbm SKV_FA_ORDER=$o
  run ord${o}_b1_host 1 32 host SKV_FA_ORDER=$o
done
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/r3_ord*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2))
    except Exception as e:
        print(f, "failed", e); print(open(f[:-4] + "err").read()[-800:])
PY
