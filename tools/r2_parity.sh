#!/usr/bin/env bash
# Round-2 parity run: (re)generate the step-level golden vectors from the reference's own CUDA path, then the GPU suite.
mkdir -p gpurun_out
if [ "$1" = "gen" ]; then
  timeout 900 python tests/golden/make_step_golden.py gpurun_out/step_golden.npz > gpurun_out/make_step_golden.log 2>&1; echo "make_step_golden rc=$?"
  tail -15 gpurun_out/make_step_golden.log
  cp gpurun_out/step_golden.npz tests/golden/step_golden.npz 2>/dev/null
fi
timeout 1200 python -m pytest tests/test_step_golden.py tests/test_gpu_reference_python.py tests/test_ref_golden.py -x -q -s -m gpu --no-header -p no:cacheprovider > gpurun_out/r2_parity_new.log 2>&1; echo "new parity tests rc=$?"
grep -v "^$" gpurun_out/r2_parity_new.log | tail -40
timeout 1800 python -m pytest tests/ -q -s -m gpu --no-header -p no:cacheprovider > gpurun_out/all_gpu_tests.log 2>&1; echo "gpu suite rc=$?"
grep -E "chunks selected differently|passed|failed|error|Error" gpurun_out/all_gpu_tests.log | tail -40
