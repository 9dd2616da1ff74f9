#!/usr/bin/env bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_engine.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/r2_headsplit.py > gpurun_out/r2_headsplit_n$N.log 2>&1; echo "headsplit check N=$N rc=$?"
grep -E "rank|Error|error|assert" gpurun_out/r2_headsplit_n$N.log | tail -12
bash tools/r2_multi.sh $N
