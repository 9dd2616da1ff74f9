#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/test_gpu_dense.log 2>&1; echo "dense rc=$?"; tail -2 gpurun_out/test_gpu_dense.log
for b in 8 32; do
  echo "== batch $b (4 layers)"
  timeout 400 python tools/bench_decode_model.py --steps 16 --layers 4 --batch $b 2>gpurun_out/dm_b$b.err > gpurun_out/decode_model_b$b.json; tail -2 gpurun_out/dm_b$b.err
  python -c "
import json
d=json.load(open('gpurun_out/decode_model_b$b.json'))
print('native %.3f ms lib %.3f ms hot %.3f' % (d['native_ms_per_token'], d['library_ms_per_token'], d['hot_ms_per_token']))
for k,v in d['ops'].items(): print(k, {a:round(x,1) for a,x in v.items()})"
done
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'])
w=d['whole_model']
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in w.items() if k not in('what','ops','note')})
for k,v in w.get('ops',{}).items(): print(k,{a:round(b,2) for a,b in v.items()})
PY
