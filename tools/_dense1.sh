#!/usr/bin/env bash
mkdir -p gpurun_out
echo skip golden

timeout 900 python -m pytest tests/test_gpu_dense.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/test_gpu_dense.log 2>&1
rc=$?; echo "dense tests rc=$rc"; tail -30 gpurun_out/test_gpu_dense.log
if [ $rc -eq 0 ]; then
  timeout 600 python tools/bench_decode_model.py --steps 32 > gpurun_out/decode_model.json 2> gpurun_out/decode_model.err; echo "bench rc=$?"
  tail -5 gpurun_out/decode_model.err; cat gpurun_out/decode_model.json
fi
