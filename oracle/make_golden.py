"""Generate tests/golden/*.pt by executing the UNMODIFIED reference under the helion stand-in.

TEST INFRASTRUCTURE.  Run in the build container only (`/root/reference` does not exist on the
GPU box):   python oracle/make_golden.py
Each fixture holds the seeded inputs and the reference's own outputs; tests/test_oracle_golden.py
asserts torch.equal between them and the `literal` restatements in oracle/.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QA_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(HERE, "_helion_standin"))
    sys.path.insert(0, REF)
    import attention_bf16 as RB
    import attention_int8 as RI
    import attention_jvp as RJ
    from helion import _state

    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # deterministic fp32 summation order inside torch.matmul

    # ---- int8 forward (+ literal backward where the reference can run it: Bq == Bkv) ----
    for (B, H, S, D, bq, bkv, seed) in [(1, 2, 128, 64, 32, 32, 11), (1, 2, 256, 64, 128, 128, 12),
                                        (2, 1, 256, 128, 128, 128, 13), (1, 2, 256, 128, 64, 128, 14)]:
        g = torch.Generator().manual_seed(seed)
        q, k, v, dO = [torch.randn(B, H, S, D, generator=g).to(torch.float16) for _ in range(4)]
        if seed == 13:  # exercise smoothing-like offsets and a zero-free but tiny block
            k = (k.float() + 3.0).to(torch.float16)
            v[:, :, :128] *= 1e-3
        _state.tunable_overrides = {"Bq": bq, "Bkv": bkv}
        out = RI.helion_atten_int8_hl_dot_fwd(q, k, v)
        fx = {"q": q, "k": k, "v": v, "Bq": bq, "Bkv": bkv,
              "fwd": [o for o in out[:8]]}
        if bq == bkv:
            kmean = (torch.randn(B, H, S, generator=g) * 0.1).to(torch.float16)
            bw = RI.helion_atten_int8_hl_dot_bwd(dO, out[2], out[5], out[3], kmean, out[6],
                                                 out[4], out[7], out[0], out[1], bq, bkv)
            fx.update({"dO": dO, "k_mean_bhk": kmean, "bwd": list(bw)})
        _state.tunable_overrides = {}
        torch.save(fx, os.path.join(OUT, f"int8_B{B}H{H}S{S}D{D}_bq{bq}_bkv{bkv}.pt"))

    # ---- bf16 forward/backward ----
    for (B, H, S, D, causal, tk, seed) in [(1, 2, 128, 64, False, 32, 21), (1, 2, 128, 64, True, 32, 22),
                                           (1, 2, 256, 128, True, 128, 23), (2, 1, 256, 128, False, 128, 24)]:
        g = torch.Generator().manual_seed(seed)
        q, k, v, dO = [torch.randn(B, H, S, D, generator=g) for _ in range(4)]
        q, k, v = q.to(torch.float16), k.to(torch.float16), v.to(torch.bfloat16)
        _state.override_block_sizes = [1, 32, tk]
        O, lse = RB.helion_atten_bf16_fwd_training(q, k, v, causal)
        _state.override_block_sizes = None          # backward uses its own Config([2,16,16])
        dq, dk, dv = RB.helion_flash_atten_2_algo_4_bwd(q, k, v, O, lse, causal, dO)
        torch.save({"q": q, "k": k, "v": v, "dO": dO, "causal": causal, "tile_k": tk,
                    "O": O, "lse": lse, "dq": dq, "dk": dk, "dv": dv},
                   os.path.join(OUT, f"bf16_B{B}H{H}S{S}D{D}_c{int(causal)}_tk{tk}.pt"))

    # ---- JVP ----
    for (B, H, S, D, ones, seed) in [(1, 2, 128, 64, True, 31), (2, 1, 128, 64, False, 32)]:
        g = torch.Generator().manual_seed(seed)
        q, k, v = [torch.randn(B, H, S, D, generator=g) for _ in range(3)]
        if ones:   # the reference test's tangents (attention_jvp.py:242-245)
            tq, tk_, tv = [torch.ones(B, H, S, D) for _ in range(3)]
        else:
            tq, tk_, tv = [torch.randn(B, H, S, D, generator=g) for _ in range(3)]
        O, tO, lse = RJ.helion_attention_jvp_forward_fp32(q, k, v, tq, tk_, tv)
        torch.save({"q": q, "k": k, "v": v, "tq": tq, "tk": tk_, "tv": tv, "O": O, "tO": tO, "lse": lse},
                   os.path.join(OUT, f"jvp_B{B}H{H}S{S}D{D}_ones{int(ones)}.pt"))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
