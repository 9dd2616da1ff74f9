#!/usr/bin/env bash
mkdir -p gpurun_out
SKV_LIB=shadowkv_b200/lib/libshadowkv_b200_trace.so timeout 300 python tools/trace_fused.py --hbm > gpurun_out/r2_trace_fused_hbm.txt 2>&1; echo "trace hbm rc=$?"
cut -c1-190 gpurun_out/r2_trace_fused_hbm.txt | head -28
SKV_LIB=shadowkv_b200/lib/libshadowkv_b200_trace.so timeout 300 python tools/trace_fused.py > gpurun_out/r2_trace_fused_host.txt 2>&1; echo "trace host rc=$?"
cut -c1-190 gpurun_out/r2_trace_fused_host.txt | head -14
