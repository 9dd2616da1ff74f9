#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 120 python tools/exp_prefetch.py 2>&1 | tail -5
for cfg in "0 5 1" "16 5 1" "64 5 1" "0 11 1" "0 5 2" "0 3 1"; do
  set -- $cfg
  echo "== L2_PREFETCH=$1 STAGES=$2 CTAS_PER_SM=$3"
  SKV_DENSE_L2_PREFETCH=$1 SKV_DENSE_STAGES=$2 SKV_DENSE_CTAS_PER_SM=$3 timeout 300 python tools/bench_decode_model.py --steps 16 --layers 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('native %.3f ms lib %.3f ms hot %.3f' % (d['native_ms_per_token'], d['library_ms_per_token'], d['hot_path_ms_per_token']))
print(' '.join('%s=%.1f' % (k, v['us']) for k, v in d['ops'].items()))"
done
