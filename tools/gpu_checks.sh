#!/usr/bin/env bash
# One gpurun call that re-validates everything measured in DESIGN.md / profiles/ (about 5 GPU-minutes):
#   gpurun --timeout 3000 -- 'bash tools/gpu_checks.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
# the whole GPU suite in ONE process, exactly as the round-end driver runs it
timeout 1500 python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1
echo "gpu tests (single process) rc=$?"; tail -3 gpurun_out/all_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
# launch list of the library's own kernels over the same command (cold caches, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k 'regex:score_|select_|fetch_|rebuild_|attention_|compact_|vstore|append_|advance' -c 1200 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-kernel-table --no-cpu-baseline \
  --hbm-store-ways 0 --no-whole-model > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
w = d.get("whole_model") or {}
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "roofline", round(d["roofline"]["frac"], 3),
      "store", round((d.get("hbm_chunk_store") or {}).get("value", 0), 1),
      "whole model native/library", round(w.get("native_tokens_per_s", 0), 1), round(w.get("library_tokens_per_s", 0), 1))
PY
