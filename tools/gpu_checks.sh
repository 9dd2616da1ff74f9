#!/usr/bin/env bash
# One gpurun call that re-validates everything measured in DESIGN.md / profiles/ (about 4 GPU-minutes):
#   gpurun --timeout 3000 -- 'bash tools/gpu_checks.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
# the whole GPU suite in ONE process, exactly as the round-end driver runs it
timeout 1800 python -m pytest tests/ -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1
echo "gpu tests (single process) rc=$?"; tail -3 gpurun_out/all_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_driver_args.json 2> gpurun_out/bench_driver_args.err; echo "bench (driver args) rc=$?"
# launch list of the library's own kernels over the same command (cold caches, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k 'regex:score_|select_|fetch_|rebuild_|attention_|compact_|vstore|append_|advance' -c 1200 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-kernel-table --no-cpu-baseline \
  --no-hbm-first --no-hit-sweep --no-whole-model --no-reference-cuda > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
# one full-set capture of one launch of each product kernel of the step (steady state: the seventh token, so the chunk cache and the threshold hints are warm)
timeout 900 ncu --set full --clock-control none --import-source on \
  -k 'regex:score_tcgen05|select_kernel|fetch_rebuild_attend' -s 576 -c 3 -f -o gpurun_out/r02_full \
  python bench.py --steps 2 --warmup 3 --no-kernel-table --no-cpu-baseline --no-hbm-first --no-hit-sweep --no-whole-model \
  --no-reference-cuda > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/r02_full.ncu-rep
python - <<'PY'
import json
for name in ("bench_default", "bench_driver_args"):
    try:
        d = json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(name, "failed:", e); print(open(f"gpurun_out/{name}.err").read()[-2000:]); continue
    w = d.get("whole_model") or {}
    rc = (d.get("cpu_baseline") or {}).get("reference_cuda") or {}
    print(name, "value", round(d["value"], 1), "us/layer", round(d["us_per_layer_step"], 2), "e2e", round(d["e2e"]["value"], 1),
          "| by hit rate", [(r["hit_rate"], round(r["value"], 1)) for r in d["e2e"]["by_hit_rate"]["rows"]],
          "| hbm_first", round((d.get("hbm_first") or {}).get("value", 0), 1), (d.get("hbm_first") or {}).get("policy"),
          "| roofline", d["roofline"]["bound"], round(d["roofline"]["frac"], 3), "hbm", round(d["roofline"]["hbm"]["frac"], 3),
          "| step", {k: round(v, 3) for k, v in d["roofline"].get("step", {}).items() if isinstance(v, float)},
          "| ref_cuda", rc.get("value"), rc.get("unavailable"), "| cpu", (d.get("cpu_baseline") or {}).get("value"),
          "| clocks", d["clocks"], "| whole model native/library", round(w.get("native_tokens_per_s", 0), 1), round(w.get("library_tokens_per_s", 0), 1),
          "|", d["active_config"])
PY
