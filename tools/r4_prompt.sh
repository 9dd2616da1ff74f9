#!/usr/bin/env bash
# Round-2 (session 4): prompt path of the model mirror (prefill -> decode) -- tests, then the e2e-shaped run.
# SKIP4=1 skips the 4-layer trial run.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model_prefill.py -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/r4_prompt_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r4_prompt_tests.log
if [ -z "$SKIP4" ]; then
  timeout 600 python tools/run_from_prompt.py --layers 4 --gen-len 20 > gpurun_out/r4_prompt_4layers.json 2> gpurun_out/r4_prompt_4layers.err; echo "4 layers rc=$?"; tail -3 gpurun_out/r4_prompt_4layers.err; cat gpurun_out/r4_prompt_4layers.json
fi
timeout 900 python tools/run_from_prompt.py --gen-len 90 > gpurun_out/r4_prompt_llama31_122k_b1.json 2> gpurun_out/r4_prompt_full.err; echo "full rc=$?"; tail -3 gpurun_out/r4_prompt_full.err; cat gpurun_out/r4_prompt_llama31_122k_b1.json
timeout 900 python tools/run_from_prompt.py --gen-len 90 --v-residency hbm > gpurun_out/r4_prompt_llama31_122k_b1_hbm.json 2> gpurun_out/r4_prompt_full_hbm.err; echo "full hbm rc=$?"; tail -3 gpurun_out/r4_prompt_full_hbm.err; cat gpurun_out/r4_prompt_llama31_122k_b1_hbm.json
