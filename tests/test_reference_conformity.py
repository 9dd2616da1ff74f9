"""Drop-in conformity of the Python shim against the reference tree itself (runs only where /root/reference is
mounted, i.e. in the build container; skipped on the GPU box): every function the reference binds with pybind11
(kernels/main.cu) exists in shadowkv_b200.shadowkv with the same number of positional parameters as its C++
declaration (kernels/functions.h), and the cache classes expose every method models/base.py calls on them."""
import inspect
import os
import re

import pytest

REF = os.environ.get("SHADOWKV_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "kernels")), reason="reference tree not present")


def _reference_functions():
    main = open(os.path.join(REF, "kernels", "main.cu"), encoding="utf-8").read()
    names = re.findall(r'm\.def\("([a-z0-9_]+)"', main)
    hdr = open(os.path.join(REF, "kernels", "functions.h"), encoding="utf-8").read()
    hdr = re.sub(r"//[^\n]*", "", hdr)
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    arity = {}
    for n in names:
        m = re.search(r"\bvoid\s+" + n + r"\s*\((.*?)\)\s*;", hdr, flags=re.S)
        assert m, f"declaration of {n} not found in functions.h"
        params = [p for p in m.group(1).split(",") if p.strip()]
        arity[n] = len(params)
    return names, arity


def test_shim_matches_the_pybind_module():
    from shadowkv_b200 import shadowkv
    names, arity = _reference_functions()
    assert sorted(names) == sorted(shadowkv.__all__)
    for n in names:
        sig = inspect.signature(getattr(shadowkv, n))
        fixed = [p for p in sig.parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        variadic = any(p.kind == p.VAR_POSITIONAL for p in sig.parameters.values())
        if variadic:      # the three push-cache wrappers take the reference's 19 trailing ints as *ints
            assert len(fixed) <= arity[n]
        else:
            assert len(fixed) == arity[n], f"{n}: reference takes {arity[n]} arguments, shim takes {len(fixed)}"


def test_cache_classes_expose_what_base_py_calls():
    from shadowkv_b200 import kv_cache
    base = open(os.path.join(REF, "models", "base.py"), encoding="utf-8").read()
    called = set(re.findall(r"self\.kv_cache\.([a-zA-Z_0-9]+)", base))
    assert {"get_retrieval_position_ids", "get_value_cache", "get_key_cache", "update_kv_cache"} <= called
    for cls in (kv_cache.ShadowKVCache_CPU, kv_cache.ShadowKVCache):
        for name in called:
            if name in ("copy_stream",):        # instance attribute, created in __init__
                continue
            assert hasattr(cls, name), f"{cls.__name__} lacks {name} (called by models/base.py)"
    ref_kv = open(os.path.join(REF, "models", "kv_cache.py"), encoding="utf-8").read()
    m = re.search(r"class ShadowKVCache_CPU.*?def __init__\(self,(.*?)\)\s*->", ref_kv, flags=re.S)
    args = re.sub(r"#[^\n]*", "", m.group(1))
    ref_params = [p.split(":")[0].split("=")[0].strip() for p in args.split(",") if p.strip()]
    ours = list(inspect.signature(kv_cache.ShadowKVCache_CPU.__init__).parameters)[1:]
    assert ours[:len(ref_params)] == ref_params, (ours, ref_params)


def test_tensor_op_wrappers_match():
    from shadowkv_b200 import tensor_op
    ref = open(os.path.join(REF, "models", "tensor_op.py"), encoding="utf-8").read()
    for n in ("apply_rotary_pos_emb_cuda", "apply_rotary_pos_emb_cuda_push_cache", "batch_gather_gemm_rotary_pos_emb_cuda"):
        m = re.search(r"def " + n + r"\((.*?)\):", ref, flags=re.S)
        args = re.sub(r"#[^\n]*", "", m.group(1))
        ref_params = [p.split(":")[0].split("=")[0].strip() for p in args.split(",") if p.strip()]
        assert list(inspect.signature(getattr(tensor_op, n)).parameters) == ref_params
