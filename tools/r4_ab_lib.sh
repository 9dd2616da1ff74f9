#!/usr/bin/env bash
# Same-box A/B of two builds of the library (SKV_LIB): tools/r4_ab_lib.sh <lib A> <lib B>
mkdir -p gpurun_out
F="--steps 64 --warmup 6 --no-hit-sweep --no-head-split --no-whole-model --no-cpu-baseline --no-reference-cuda"
for i in 1 2 3; do
  for tag in A B; do
    if [ $tag = A ]; then L=$1; else L=$2; fi
    SKV_LIB=$L timeout 600 python bench.py $F > gpurun_out/r4_ab_${tag}$i.json 2> gpurun_out/r4_ab_${tag}$i.err; echo "$tag$i rc=$?"
  done
done
python - <<'PY'
import json
for tag in ("A", "B"):
    v, h = [], []
    for i in (1, 2, 3):
        try:
            d = json.loads(open(f"gpurun_out/r4_ab_{tag}{i}.json").read().strip().splitlines()[-1])
            v.append(round(d["value"], 1)); h.append(round((d.get("hbm_first") or {}).get("value", 0), 1))
        except Exception as e:
            v.append(str(e)[:40])
    print(tag, "value", v, "hbm_first", h)
PY
