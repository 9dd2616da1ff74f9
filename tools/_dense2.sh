#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/test_gpu_dense.log 2>&1
rc=$?; echo "dense tests rc=$rc"; tail -5 gpurun_out/test_gpu_dense.log
for w in o gu; do timeout 120 python tools/trace_dense.py $w > gpurun_out/trace_dense_$w.txt 2>&1; tail -22 gpurun_out/trace_dense_$w.txt; done
timeout 600 python tools/bench_decode_model.py --steps 32 > gpurun_out/decode_model.json 2> gpurun_out/decode_model.err; echo "bench rc=$?"
tail -3 gpurun_out/decode_model.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/decode_model.json'))
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k not in('what','ops')})
for k,v in d.get('ops',{}).items(): print(k,{a:round(b,2) for a,b in v.items()})
PY
