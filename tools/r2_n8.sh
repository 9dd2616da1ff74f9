#!/usr/bin/env bash
bash tools/r2_multi.sh 8
cp gpurun_out/r2_bench_n8.json gpurun_out/r2_default_n8.json
bash tools/r2_cfg.sh 8 cfg4_glm_128k_b16_n8 --model glm-4-9b-1m --ctx 131072 --batch 2
bash tools/r2_cfg.sh 8 cfg5_yi_200k_b32_n8 --model yi-9b-200k --ctx 204800 --batch 4 --v-residency hbm
