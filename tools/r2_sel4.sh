#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_score.py -x -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -2
bash tools/r2_batch.sh 8 2>&1 | grep -v "^  h\|^  attention\|^  score_select"
bash tools/r2_batch.sh 1 2>&1 | grep -v "^  h\|^  attention\|^  score_select"
