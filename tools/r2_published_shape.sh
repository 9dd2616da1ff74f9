#!/usr/bin/env bash
# the reference's published throughput shape (test/e2e.py:161-166: Llama-3.1-8B, 122K, gen_len 100) as a
# whole-model run on one B200, at the largest batches the box holds per V residency tier.
mkdir -p gpurun_out
free -g | head -2
timeout 1200 python tools/bench_decode_model.py --batch ${1:-12} --steps 100 --warmup 4 --skip-ops --v-residency hbm --modes hot,native \
  > gpurun_out/r2_published_b${1:-12}_hbm.json 2> gpurun_out/r2_published_b${1:-12}_hbm.err; echo "batch ${1:-12} hbm rc=$?"
tail -c 600 gpurun_out/r2_published_b${1:-12}_hbm.err
if [ -n "$2" ]; then
  timeout 1500 python tools/bench_decode_model.py --batch $2 --steps 100 --warmup 4 --skip-ops --v-residency host --modes hot,native \
    > gpurun_out/r2_published_b$2_host.json 2> gpurun_out/r2_published_b$2_host.err; echo "batch $2 host rc=$?"
  tail -c 600 gpurun_out/r2_published_b$2_host.err
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_published_b*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "batch", d["batch"], d["v_residency"], "native tok/s", round(d["native_tokens_per_s"], 1), "ms/token", round(d["native_ms_per_token"], 3),
              "hot tok/s", round(d["hot_tokens_per_s"], 1), "hit", d["chunk_cache_hit_rate"], "V GB", round(d["v_store_GB"], 1), "setup", d["setup_s"])
    except Exception as e:
        print(f, "failed", e)
PY
