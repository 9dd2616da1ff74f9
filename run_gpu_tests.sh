#!/usr/bin/env bash
# Runs the GPU test files one by one (own timeout each) and keeps their logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
rc_all=0
for f in test_gpu_integer test_gpu_rope test_gpu_score test_gpu_rebuild test_gpu_attention test_gpu_decode test_gpu_engine test_gpu_prefill test_gpu_dense test_ref_golden; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/$f.log 2>&1
  rc=$?
  echo "$f rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$f.log
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
