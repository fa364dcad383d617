"""CPU checks of the NVFP4 definition (oracle/fp4_ref.py).  The reference ships no FP4 code (README.md:48-54 only names it), so
the oracle cannot be pinned against reference outputs; these tests pin its building blocks against independent brute-force
statements instead: e2m1 round-to-nearest-even, the tcgen05.cp scale-factor atom layout, step-size independence of the online
softmax, and agreement with the fp32 attention math."""
import itertools

import torch

from oracle import fp4_ref
from oracle.baseline import baseline_pytorch_attention

GRID = [0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0]


def _brute_e2m1(y: float) -> float:
    a = min(abs(y), 6.0)
    best = min(range(8), key=lambda i: (abs(GRID[i] - a), i & 1))       # nearest; on a tie the even code
    return -GRID[best] if y < 0 else GRID[best]


def test_e2m1_round_to_nearest_even_matches_brute_force():
    mids = [(GRID[i] + GRID[i + 1]) / 2 for i in range(7)]
    pts = sorted(set(GRID + mids + [m + d for m in mids for d in (-1e-6, 1e-6)] + [6.5, 7.0, 100.0, 1e-9]))
    ys = torch.tensor([s * p for p, s in itertools.product(pts, (1.0, -1.0))], dtype=torch.float32)
    vals, codes = fp4_ref.e2m1_rn(ys)
    for y, v, c in zip(ys.tolist(), vals.tolist(), codes.tolist()):
        assert v == _brute_e2m1(y), (y, v)
        assert GRID[c & 7] == abs(v) and bool(c & 8) == (y < 0 or (y == 0 and str(y).startswith("-")))


def test_scale_factor_atom_layout():
    sf = (torch.arange(256 * 8) % 251).to(torch.uint8).reshape(256, 8)           # two 128-row tiles, D / 16 = 8 blocks
    atoms = fp4_ref.sf_atoms(sf)                                                  # [tile, k step, 512]
    assert atoms.shape == (2, 2, 512)
    for t, k, r, s in itertools.product(range(2), range(2), (0, 1, 31, 32, 77, 127), range(4)):
        assert atoms[t, k, 16 * (r % 32) + 4 * (r // 32) + s] == sf[t * 128 + r, 4 * k + s]


def test_two_level_quantisation_is_consistent():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 64, 128, generator=g) * torch.tensor([1e-3, 1.0, 50.0]).view(3, 1, 1)
    x[1, :, 16:32] = 0
    deq, codes, sfb, sg = fp4_ref.quant_nvfp4(x)
    assert torch.all(sg == x.abs().amax(dim=(1, 2)) / 2688.0)
    assert int(codes.max()) < 16 and torch.all(codes[1, :, 16:32] == 0) and torch.all(sfb[1, :, 1] == 0)
    blk_amax = x.reshape(3, 64, 8, 16).abs().amax(-1, keepdim=True)
    err = (deq - x).reshape(3, 64, 8, 16).abs()
    assert torch.all(err <= 0.27 * blk_amax + 1e-12)                              # half an e2m1 step at the top of the block (1 of 6) + scale rounding


def test_forward_is_step_size_independent_and_close_to_fp32_math():
    g = torch.Generator().manual_seed(2)
    q, k, v = [torch.randn(1, 2, 384, 128, generator=g).half() for _ in range(3)]
    O128, lse128, _ = fp4_ref.fp4_fwd(q, k, v, step=128)
    O64, lse64, _ = fp4_ref.fp4_fwd(q, k, v, step=64)
    assert (lse128 - lse64).abs().max() < 1e-4                                     # l sums the unquantised P: identical up to fp32 order
    cos = torch.nn.functional.cosine_similarity(O128.float().flatten(), O64.float().flatten(), dim=0)
    assert cos > 0.999                                                             # P is re-quantised against a different running maximum
    base = baseline_pytorch_attention(q.float(), k.float(), v.float(), 128, False)
    assert torch.nn.functional.cosine_similarity(O128.float().flatten(), base.flatten(), dim=0) > 0.975


def test_ragged_and_causal_follow_the_masked_definition():
    g = torch.Generator().manual_seed(3)
    q, k, v = [torch.randn(1, 1, 200, 128, generator=g).half() for _ in range(3)]
    for causal in (False, True):
        O, lse, qi = fp4_ref.fp4_fwd(q, k, v, causal=causal)
        assert O.shape == (1, 1, 200, 128) and lse.shape == (1, 200) and torch.isfinite(O.float()).all()
        assert qi["q4"].shape == (256, 64) and torch.all(qi["k4"][200:] == 0)     # padded rows: zero codes (K stays zero after smoothing)
        base = baseline_pytorch_attention(q.float(), k.float(), v.float(), 128, causal)
        assert torch.nn.functional.cosine_similarity(O.float().flatten(), base.flatten(), dim=0) > 0.97
