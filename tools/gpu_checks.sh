#!/usr/bin/env bash
# One gpurun call that re-validates everything measured in DESIGN.md / profiles/ (about 3 GPU-minutes):
#   gpurun --timeout 3000 -- 'bash tools/gpu_checks.sh'
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
bash run_gpu_tests.sh > gpurun_out/all_tests.log 2>&1; echo "gpu tests rc=$?"; cat gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/test_gpu_dense.log 2>&1
echo "dense rc=$?"; tail -2 gpurun_out/test_gpu_dense.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
# launch list of the library's own kernels over the same command (cold caches, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k 'regex:score_|select_|fetch_|rebuild_|attention_|dense_|rmsnorm|argmax|l2_prefetch|compact_|vstore|append_|advance' -c 3000 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-kernel-table --no-cpu-baseline \
  --hbm-store-ways 0 --whole-model-steps 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
