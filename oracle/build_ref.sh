#!/usr/bin/env bash
# Builds the UNMODIFIED reference extension (reference setup.py:28-55 lists the
# six translation units) for sm_100a straight from /root/reference/kernels into
# oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun).
# Nothing is copied from the reference; objects and the .so are the only outputs.
# Test infrastructure: used to pin the oracle and as the "R-CUDA" timing arm.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SHADOWKV_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"
SITE="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["purelib"])')"
CUT="$SITE/flashinfer/data/cutlass"
TORCH="$SITE/torch"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
mkdir -p "$OUT/kernels"
if [ ! -d "$REF/kernels" ]; then echo "reference not present at $REF; keeping prebuilt files" ; exit 0; fi
# The reference's own Python for this path (models/kv_cache.py: ShadowKVCache_CPU; models/tensor_op.py: the
# stride-marshalling wrappers), verbatim, so tests can run it UNMODIFIED on top of either native module
# (tests/refpy.py).  oracle/_ref is git-ignored: nothing of the reference enters the history.
mkdir -p "$OUT/models"
cp "$REF/models/kv_cache.py" "$REF/models/tensor_op.py" "$OUT/models/"
: > "$OUT/models/__init__.py"
pids=()
for f in main rope rope_new gather_copy batch_gather_gemm batch_gemm_softmax; do
  if [ "$OUT/$f.o" -nt "$REF/kernels/$f.cu" ]; then continue; fi
  nvcc -std=c++17 --expt-relaxed-constexpr -O3 \
       -gencode arch=compute_100a,code=sm_100a \
       -I"$CUT/include" -I"$CUT/tools/util/include" -I"$HERE/ref_shim" -I"$REF/kernels" \
       -I"$PYINC" -I"$TORCH/include" -I"$TORCH/include/torch/csrc/api/include" \
       -DTORCH_EXTENSION_NAME=shadowkv -D_GLIBCXX_USE_CXX11_ABI=1 -Xcompiler -fPIC \
       -c "$REF/kernels/$f.cu" -o "$OUT/$f.o" > "$OUT/$f.log" 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]:-}"; do [ -z "$p" ] || wait "$p" || rc=1; done
[ $rc -eq 0 ] || { echo "reference compile failed; see $OUT/*.log"; exit 1; }
g++ -shared -o "$OUT/kernels/shadowkv.cpython-312-x86_64-linux-gnu.so" "$OUT"/*.o \
    -L"$TORCH/lib" -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda \
    -L/usr/local/cuda/lib64 -lcudart
echo "built $OUT/kernels/shadowkv.cpython-312-x86_64-linux-gnu.so"
