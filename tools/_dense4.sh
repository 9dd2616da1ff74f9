#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/test_gpu_dense.log 2>&1
rc=$?; echo "dense tests rc=$rc"; tail -5 gpurun_out/test_gpu_dense.log
for w in o gu; do timeout 120 python tools/trace_dense.py $w > gpurun_out/trace_dense_$w.txt 2>&1; tail -11 gpurun_out/trace_dense_$w.txt; done
for pf in 0 48 96; do
  echo "== l2 prefetch $pf MB"
  timeout 300 python tools/bench_decode_model.py --steps 16 --layers 8 --l2-prefetch-mb $pf 2>gpurun_out/decode_model.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('native %.3f ms lib %.3f ms hot %.3f' % (d['native_ms_per_token'], d['library_ms_per_token'], d['hot_path_ms_per_token']))
print(' '.join('%s=%.1f' % (k, v['us']) for k, v in d['ops'].items()))"
  tail -2 gpurun_out/decode_model.err
done
