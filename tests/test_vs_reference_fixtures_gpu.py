"""CUDA kernels against the outputs of the UNMODIFIED reference (tests/golden/*.pt, produced by oracle/make_golden.py
by executing /root/reference/*.py under the eager helion stand-in).  Tolerances as stated per path."""
import glob
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_bf16_fwd_vs_reference_outputs(golden_dir):
    from quantizedattention_b200 import attention_bf16 as A
    n = 0
    for f in sorted(glob.glob(os.path.join(golden_dir, "bf16_*_c0_*.pt"))):      # non-causal: literal == contract up to rounding
        fx = torch.load(f, weights_only=True)
        O, lse = A.helion_atten_bf16_fwd_training(fx["q"].cuda(), fx["k"].cuda(), fx["v"].cuda(), False)
        assert (O.cpu() - fx["O"]).abs().max() < 2.5e-2, f
        assert (lse.cpu() - fx["lse"]).abs().max() < 5e-2, f      # literal m ratchets (LEDGER B-2): lse equal in value, not in (m, l) split
        n += 1
    assert n == 2


def test_bf16_fwd_causal_vs_reference_outputs(golden_dir):
    """Causal: the literal's finite -126 fill leaks weight into masked keys for the first rows (LEDGER B-3); rows >= 64
    agree to tolerance, row 0 (uniform over all keys) agrees exactly in both."""
    from quantizedattention_b200 import attention_bf16 as A
    for f in sorted(glob.glob(os.path.join(golden_dir, "bf16_*_c1_*.pt"))):
        fx = torch.load(f, weights_only=True)
        O, _ = A.helion_atten_bf16_fwd_training(fx["q"].cuda(), fx["k"].cuda(), fx["v"].cuda(), True)
        assert (O.cpu()[:, :, 64:] - fx["O"][:, :, 64:]).abs().max() < 3e-2, f
        assert (O.cpu()[:, :, 0] - fx["O"][:, :, 0]).abs().max() < 1e-2, f


def test_jvp_vs_reference_outputs(golden_dir):
    from quantizedattention_b200 import attention_jvp as J
    for f in sorted(glob.glob(os.path.join(golden_dir, "jvp_*.pt"))):
        fx = torch.load(f, weights_only=True)
        O, tO, lse = J.helion_attention_jvp_forward_fp32(*[fx[n].cuda() for n in ("q", "k", "v", "tq", "tk", "tv")])
        assert (O.cpu() - fx["O"]).abs().max() < 1e-2, f            # the reference's own atol (attention_jvp.py:260)
        assert (tO.cpu() - fx["tO"]).abs().max() < 4e-2, f
        assert (lse.cpu() - fx["lse"]).abs().max() < 5e-3, f


def test_int8_quantised_tensors_vs_reference_outputs(golden_dir):
    """All four int8 fixtures: q/k/v int8 tensors and fp16 scales are bit-identical to the reference's."""
    from quantizedattention_b200 import ops
    for f in sorted(glob.glob(os.path.join(golden_dir, "int8_*.pt"))):
        fx = torch.load(f, weights_only=True)
        O_ref, lse_ref, q_i8, k_i8_T, v_i8, sq, sk, sv = fx["fwd"]
        for x, blk, ref_i8, ref_s in ((fx["q"], fx["Bq"], q_i8, sq), (fx["k"], fx["Bkv"], k_i8_T.t(), sk), (fx["v"], fx["Bkv"], v_i8, sv)):
            qi, s = ops.quant_block(x.cuda(), blk)
            assert torch.equal(qi.cpu(), ref_i8) and torch.equal(s.cpu(), ref_s), f
