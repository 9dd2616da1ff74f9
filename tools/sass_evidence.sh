#!/usr/bin/env bash
# SASS evidence per object file (run HERE after `make`; no GPU needed): which tensor-core, TMEM, TMA, bulk-copy, cluster
# and warp-match instructions each translation unit of libshadowkv_b200.so contains.  Mnemonics as listed in
# /opt/skills/guides/B200_PROFILING.md: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA),
# UBLKCP = cp.async.bulk, HMMA = mma.sync, LDGSTS = cp.async, UCGABAR = barrier.cluster, MATCH = match.any.
echo "# cuobjdump -sass build/*.o, instruction counts per translation unit (nvcc $(nvcc --version | grep -o 'release [0-9.]*'), -gencode arch=compute_100a,code=sm_100a)"
for f in score_mma score rebuild fused_attend gather attention_mma dense prefill score_select; do
  echo; echo "== shadowkv_b200/csrc/$f.cu"
  cuobjdump -sass build/$f.o | grep -oE "\b(UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG[.0-9A-Z_]*|UTMASTG[.0-9A-Z_]*|UBLKCP[.A-Z0-9_]*|HMMA[.0-9A-Z_]*|UTCATOMSWS[.A-Z0-9_]*|LDGSTS[.A-Z0-9_]*|UCGABAR_ARV|UCGABAR_WAIT|SYNCS[.A-Z0-9_]*|MATCH[.A-Z0-9_]*|HFMA2[.A-Z0-9_]*BF16[.A-Z0-9_]*|ACQBULK)" | sort | uniq -c | sort -rn
  echo "-- first occurrences (function : instruction)"
  cuobjdump -sass build/$f.o | awk '/Function :/ {fn=$3} /UTCHMMA|LDTM|UTMALDG|UBLKCP|HMMA\.16816/ {if (!seen[fn $2]++ && n[fn]++ < 3) print "   " fn " : " $0}' | cut -c1-200 | head -12
done
