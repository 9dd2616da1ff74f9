#!/usr/bin/env bash
mkdir -p gpurun_out
for cfg in "1 11" "0 11" "1 5" "0 5" "1 8" "0 8" "1 7" "1 6"; do
  set -- $cfg
  echo "== ORDER=$1 STAGES=$2"
  SKV_DENSE_ORDER=$1 SKV_DENSE_STAGES=$2 timeout 300 python tools/bench_decode_model.py --steps 16 --layers 8 --l2-prefetch-mb 96 2>gpurun_out/decode_model.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('native %.3f ms lib %.3f ms hot %.3f' % (d['native_ms_per_token'], d['library_ms_per_token'], d['hot_path_ms_per_token']))
print(' '.join('%s=%.1f' % (k, v['us']) for k, v in d['ops'].items()))"
done
