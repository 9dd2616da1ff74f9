#!/usr/bin/env bash
for w in o down; do timeout 120 python tools/trace_dense.py $w > gpurun_out/trace_dense_$w.txt 2>&1; tail -16 gpurun_out/trace_dense_$w.txt; done
