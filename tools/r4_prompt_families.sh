#!/usr/bin/env bash
# Prompt -> decode runs at the GLM-4-9B-1M (128K, 2 sequences) and Yi-9B-200K (200K) shapes of BASELINE configs 4 / 5.
mkdir -p gpurun_out
timeout 900 python tools/run_from_prompt.py --model glm-4-9b-1m --ctx 131072 --batch 2 --gen-len 90 --v-residency hbm > gpurun_out/r4_prompt_glm_128k_b2.json 2> gpurun_out/r4_prompt_glm.err; echo "glm rc=$?"; tail -3 gpurun_out/r4_prompt_glm.err; cat gpurun_out/r4_prompt_glm_128k_b2.json
timeout 900 python tools/run_from_prompt.py --model yi-9b-200k --ctx 204800 --batch 1 --gen-len 90 --v-residency hbm > gpurun_out/r4_prompt_yi_200k_b1.json 2> gpurun_out/r4_prompt_yi.err; echo "yi rc=$?"; tail -3 gpurun_out/r4_prompt_yi.err; cat gpurun_out/r4_prompt_yi_200k_b1.json
