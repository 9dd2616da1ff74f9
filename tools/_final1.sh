#!/usr/bin/env bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
bash run_gpu_tests.sh > gpurun_out/all_tests.log 2>&1; echo "gpu tests rc=$?"; cat gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_dense.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/test_gpu_dense.log 2>&1; echo "dense rc=$?"; tail -2 gpurun_out/test_gpu_dense.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_reference.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'roofline',d['roofline']['frac'])
w=d['whole_model']
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in w.items() if k not in('what','ops','note')})
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-kernel-table --no-cpu-baseline --hbm-store-ways 0 --whole-model-steps 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches_r1_final.csv
