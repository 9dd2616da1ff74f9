#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 64 > gpurun_out/bench_wm.json 2> gpurun_out/bench_wm.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_wm.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_wm.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'])
w=d['whole_model']
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in w.items() if k not in('what','ops','note')})
for k,v in w.get('ops',{}).items(): print(k,{a:round(b,2) for a,b in v.items()})
PY
