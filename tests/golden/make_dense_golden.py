"""Generates tests/golden/dense_golden.npz: outputs of the LIBRARY kernels the reference's dense decode layer calls
(models/llama.py:354-427) -- flashinfer rmsnorm, torch F.linear on bf16 (cuBLAS), vllm rotary_embedding and
silu_and_mul -- on small seeded inputs, run on a B200 with the builds this image carries.  Run on the GPU box:

    python tests/golden/make_dense_golden.py gpurun_out/dense_golden.npz

The fixture pins oracle/dense_oracle.py (CPU test tests/test_dense_oracle.py).  Sections whose library cannot be
imported or run are left out and listed under "missing"."""
import sys

import numpy as np
import torch

BF16 = torch.bfloat16


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int16).numpy()


def main(path):
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(20251018)
    out, missing = {}, []

    def rn(*s, scale=1.0):
        return (torch.randn(*s, generator=g) * scale).to(BF16)

    B, hid, Hq, H, ffn = 3, 1024, 4, 2, 1536
    x = rn(B, hid)
    w = (1.0 + 0.1 * torch.randn(hid, generator=g)).to(BF16)
    out["x"], out["norm_w"] = bits(x), bits(w)

    # ---- flashinfer rmsnorm (tensor_op.py:29,34-39) ----
    try:
        import flashinfer
        y = flashinfer.norm.rmsnorm(x.to(dev), w.to(dev), 1e-5)
        torch.cuda.synchronize()
        out["rmsnorm"] = bits(y)
        out["flashinfer_version"] = np.array(flashinfer.__version__)
    except Exception as e:  # noqa: BLE001
        missing.append(f"flashinfer rmsnorm: {e!r}"[:200])

    # ---- F.linear, with and without bias (llama.py:378; glm.py:344) ----
    W = rn((Hq + 2 * H) * 128, hid, scale=0.02)
    bias = rn((Hq + 2 * H) * 128, scale=0.02)
    h = rn(B, hid)
    out["lin_x"], out["lin_W"], out["lin_b"] = bits(h), bits(W), bits(bias)
    out["linear"] = bits(torch.nn.functional.linear(h.to(dev), W.to(dev)))
    out["linear_bias"] = bits(torch.nn.functional.linear(h.to(dev), W.to(dev), bias.to(dev)))
    a, r = rn(B, hid), rn(B, hid)
    out["add_a"], out["add_b"], out["add"] = bits(a), bits(r), bits(a.to(dev) + r.to(dev))
    out["logits_f32"] = torch.nn.functional.linear(h.to(dev), W.to(dev)).float().cpu().numpy()

    # ---- vllm rotary_embedding (llama.py:296 neox / glm.py:421 interleaved, rot 64) and silu_and_mul (llama.py:421) ----
    try:
        import vllm._custom_ops as ops
        pos_val = 4321
        inv = 1.0 / (500000.0 ** (torch.arange(0, 128, 2, dtype=torch.float64) / 128))
        ang = torch.arange(8192, dtype=torch.float64)[:, None] * inv[None, :]
        cs_neox = torch.cat([ang.cos(), ang.sin()], dim=-1).to(BF16)            # [rows, 128] = [cos64 | sin64]
        inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.float64) / 64))
        ang = torch.arange(8192, dtype=torch.float64)[:, None] * inv[None, :]
        cs_glm = torch.cat([ang.cos(), ang.sin()], dim=-1).to(BF16)             # [rows, 64] = [cos32 | sin32]
        q, k = rn(B, 1, Hq * 128), rn(B, 1, H * 128)
        pos = torch.full((B, 1), pos_val, dtype=torch.long, device=dev)
        out["rope_q"], out["rope_k"], out["rope_pos"] = bits(q), bits(k), np.array(pos_val)
        out["cs_neox_row"], out["cs_glm_row"] = bits(cs_neox[pos_val]), bits(cs_glm[pos_val])
        for name, cs, neox in (("neox", cs_neox, True), ("glm", cs_glm, False)):
            qd, kd = q.to(dev).clone(), k.to(dev).clone()
            ops.rotary_embedding(pos, qd, kd, 128, cs.to(dev), neox)
            torch.cuda.synchronize()
            out[f"rope_q_{name}"], out[f"rope_k_{name}"] = bits(qd), bits(kd)
        gu = rn(B, 2 * ffn, scale=2.0)
        o = torch.empty(B, ffn, dtype=BF16, device=dev)
        if hasattr(ops, "silu_and_mul"):
            ops.silu_and_mul(o, gu.to(dev))
        else:
            torch.ops._C.silu_and_mul(o, gu.to(dev))
        torch.cuda.synchronize()
        out["silu_in"], out["silu_and_mul"] = bits(gu), bits(o)
        import vllm
        out["vllm_version"] = np.array(vllm.__version__)
    except Exception as e:  # noqa: BLE001
        missing.append(f"vllm ops: {e!r}"[:200])

    out["missing"] = np.array(missing)
    out["torch_version"] = np.array(torch.__version__)
    out["device"] = np.array(torch.cuda.get_device_name(0))
    np.savez_compressed(path, **out)
    print("wrote", path, "sections:", sorted(out), "missing:", missing)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/dense_golden.npz")
