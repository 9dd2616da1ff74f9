#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_score.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_selwide_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2_selwide_tests.log
for w in 6144 1000000; do
  SKV_SELECT_WIDE2_FROM=$w bash tools/r2_cfg.sh 1 cfg3_b1_wide2_$w --model llama-3-8b-1m --ctx 524288 --batch 1 --layers 8 --no-hbm-first
done
SKV_SELECT_WIDE2_FROM=1000000 SKV_SELECT_WIDE_FROM=1024 bash tools/r2_cfg.sh 1 cfg2_wide512 --layers 8 --no-hbm-first
SKV_SELECT_WIDE2_FROM=1024 bash tools/r2_cfg.sh 1 cfg2_wide1024 --layers 8 --no-hbm-first
bash tools/r2_cfg.sh 1 cfg2_base --layers 8 --no-hbm-first
