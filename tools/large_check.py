"""Full-size sanity of the fused kernels: (1) run-to-run determinism with every SM busy over several waves (a
synchronisation bug shows up as differing bits), (2) a few heads against the CPU oracle at benchmark sequence lengths.
Usage: python tools/large_check.py            (about two minutes, most of it the CPU oracle)"""
import json
import sys

import torch

sys.path.insert(0, ".")
from oracle import bf16_ref, int8_ref, jvp_ref  # noqa: E402
from quantizedattention_b200 import ops  # noqa: E402

res = {}
torch.manual_seed(0)


def same(outs):
    return all(torch.equal(outs[0], o) for o in outs[1:])


# ---------------------------------------------------------------- int8 forward / backward, cfg3 sequence length
BH, S, D = 296, 8192, 128
q, k, v, dO = [torch.randn(BH, S, D, device="cuda", dtype=torch.float16) for _ in range(4)]
qi, sq = ops.quant_block(q, 128); ki, sk = ops.quant_block(k, 128); vi, sv = ops.quant_block(v, 128)
doi, sdo = ops.quant_block(dO, 128)
for causal in (False, True):
    outs = [ops.int8_fwd_prequant(qi, ki, vi, sq, sk, sv, BH, S, S, D, causal=causal) for _ in range(3)]
    torch.cuda.synchronize()
    res[f"int8_fwd_deterministic_causal{int(causal)}"] = same([o[0] for o in outs]) and same([o[2] for o in outs])
    O, lse16, lse32 = outs[0]
    assert torch.isfinite(O.float()).all()
    for h in (0, BH - 1):                                    # oracle on two heads (first / last wave)
        sl = slice(h * S, (h + 1) * S)
        state_in = (qi[sl].cpu(), sq[h * 64:(h + 1) * 64].cpu(), ki[sl].cpu(), vi[sl].cpu(), sk[h * 64:(h + 1) * 64].cpu(), sv[h * 64:(h + 1) * 64].cpu())
        if not causal:
            Or, l16, l32 = int8_ref.int8_attend_state(state_in[0], state_in[1], state_in[2], state_in[3], state_in[4], state_in[5],
                                                      None, 1, S, S, D, 128, 128, last=True)
            res[f"int8_fwd_head{h}_max_abs_vs_oracle"] = (O[sl].cpu().float() - Or.float()).abs().max().item()
            res[f"int8_fwd_head{h}_lse_err"] = (lse32[sl].cpu() - l32).abs().max().item()
    delta = ops.bwd_delta(dO.view(-1, D), O)
    g = [ops.int8_bwd_prequant(qi, ki, vi, doi, sq, sk, sv, sdo, lse32, delta, None, BH, S, D, causal=causal) for _ in range(2)]
    torch.cuda.synchronize()
    res[f"int8_bwd_dk_dv_deterministic_causal{int(causal)}"] = torch.equal(g[0][1], g[1][1]) and torch.equal(g[0][2], g[1][2])
    res[f"int8_bwd_dq_max_rel_run_to_run_causal{int(causal)}"] = ((g[0][0].float() - g[1][0].float()).abs().max() / g[0][0].float().abs().max()).item()
    assert all(torch.isfinite(t.float()).all() for t in g[0])
del q, k, v, dO, qi, ki, vi, doi, outs, g, O, delta
torch.cuda.empty_cache()

# ---------------------------------------------------------------- bf16 forward (two-query-tile kernel), S = 8192
B, H, S, D = 4, 37, 8192, 128
q, k = [torch.randn(B, H, S, D, device="cuda", dtype=torch.float16) for _ in range(2)]
v = torch.randn(B, H, S, D, device="cuda", dtype=torch.bfloat16)
for causal in (False, True):
    outs = [ops.bf16_fwd(q, k, v, causal) for _ in range(3)]
    torch.cuda.synchronize()
    res[f"bf16_fwd_deterministic_causal{int(causal)}"] = same([o[0] for o in outs]) and same([o[1] for o in outs])
    O, lse = outs[0]
    for (b, h) in ((0, 0), (B - 1, H - 1)):
        Or, lr = bf16_ref.bf16_fwd(q[b:b + 1, h:h + 1].cpu(), k[b:b + 1, h:h + 1].cpu(), v[b:b + 1, h:h + 1].cpu(), causal,
                                   tile_k=64, mode="contract", lazy_tau=ops.BF16_RESCALE_TAU)
        res[f"bf16_fwd_causal{int(causal)}_head{b}_{h}_max_abs_vs_oracle"] = (O[b, h].cpu() - Or[0, 0]).abs().max().item()
del q, k, v, outs
torch.cuda.empty_cache()

# ---------------------------------------------------------------- JVP, cfg4 shape family
B, H, S, D = 4, 74, 4096, 64
t = [torch.randn(B, H, S, D, device="cuda") for _ in range(6)]
outs = [ops.jvp_fwd(*t) for _ in range(3)]
torch.cuda.synchronize()
res["jvp_deterministic"] = same([o[0] for o in outs]) and same([o[1] for o in outs])
O, tO, lse = outs[0]
for (b, h) in ((0, 0), (B - 1, H - 1)):
    ref = jvp_ref.jvp_fwd(*[x[b:b + 1, h:h + 1].cpu() for x in t], tile_k=128, operand_dtype=torch.bfloat16)
    res[f"jvp_head{b}_{h}_O_err"] = (O[b, h].cpu() - ref[0][0, 0]).abs().max().item()
    res[f"jvp_head{b}_{h}_tO_err"] = (tO[b, h].cpu() - ref[1][0, 0]).abs().max().item()
print(json.dumps(res, indent=1))
