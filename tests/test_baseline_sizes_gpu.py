"""Oracle parity of the backward kernels AT the BASELINE.json sequence lengths: the small-shape
tests bound the arithmetic, these bound it where every barrier phase, TMA coordinate and causal tile skip of the
benchmark shapes is exercised.

  * int8 backward, configs[2] tile shape: S = 8192, D = 128, two heads, both kernels, against `int8_bwd_contract`
    (the oracle of reference attention_int8.py:342-428 under the 8-LEDGER contract) with the bars of the small tests.
  * bf16-path backward, configs[1]: S = 4096, D = 128, causal, a two-head slice, against `bf16_bwd(mode="contract")`
    (reference attention_bf16.py:361-444).
  * ring KV on real GPUs under torchrun (needs >= 2 devices; skipped on a single-GPU box): tools/ring_bench.py check.
"""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cos(a, b):
    return torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


@pytest.mark.parametrize("kernel", ["ws", "8warp"])
def test_int8_bwd_cfg3_sequence_length_matches_contract_oracle(kernel):
    from oracle import int8_ref
    from quantizedattention_b200 import attention_int8 as A
    torch.set_num_threads(os.cpu_count() or 1)
    shape = (1, 2, 8192, 128)
    g = torch.Generator().manual_seed(8192)
    q, k, v, dO = [torch.randn(shape, generator=g).to(torch.float16) for _ in range(4)]
    k = (k.float() + 0.5).to(torch.float16)                 # non-zero token mean: exercises the k_mean term of dQ
    out = A.SageAttention3_Int8_autograd_function.forward(q.cuda(), k.cuda(), v.cuda())
    O, lse16, kmean, q_i8, k_i8_T, v_i8, sq, sk, sv, Bq, Bkv = out
    got = A.helion_atten_int8_hl_dot_bwd(dO.cuda(), q_i8, sq, k_i8_T, kmean, sk, v_i8, sv, O, lse16, Bq, Bkv, kernel=kernel)
    torch.cuda.synchronize()
    c = lambda t: t.cpu()
    ref = int8_ref.int8_bwd_contract(dO, c(q_i8), c(sq), c(k_i8_T), c(kmean), c(sk), c(v_i8), c(sv), c(O), c(lse16), Bq, Bkv)
    for name, a, b in zip(("dq", "dk", "dv"), got, ref):
        assert torch.isfinite(a.float()).all(), name
        assert _cos(a.cpu(), b) > 0.9995 and _rel(a.cpu(), b) < 3e-2, (name, _cos(a.cpu(), b), _rel(a.cpu(), b))


def test_bf16_bwd_cfg2_causal_matches_contract_oracle():
    from oracle import bf16_ref
    from quantizedattention_b200 import attention_bf16 as A
    torch.set_num_threads(os.cpu_count() or 1)
    shape = (1, 2, 4096, 128)                              # a B*H slice of configs[1] (B=4 H=16 S=4096 D=128 causal)
    g = torch.Generator().manual_seed(4096)
    q, k, v, dO = [torch.randn(shape, generator=g) for _ in range(4)]
    q, k, v = q.half(), k.half(), v.bfloat16()
    O, lse = A.helion_atten_bf16_fwd_training(q.cuda(), k.cuda(), v.cuda(), True)
    got = A.helion_flash_atten_2_algo_4_bwd(q.cuda(), k.cuda(), v.cuda(), O, lse, True, dO.cuda())
    torch.cuda.synchronize()
    ref = bf16_ref.bf16_bwd(q, k, v, O.cpu(), lse.cpu(), True, dO, mode="contract", tile_q=512, tile_k=512)
    for name, a, b in zip(("dq", "dk", "dv"), got, ref):
        assert a.dtype == torch.float32 and torch.isfinite(a).all(), name
        assert _rel(a.cpu(), b) < 6e-3, (name, _rel(a.cpu(), b))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="the NCCL ring needs at least two GPUs on the box")
@pytest.mark.parametrize("mode", ["check", "check_causal"])
def test_ring_kv_nccl_matches_single_device(mode):
    n = 2 if torch.cuda.device_count() < 4 else 4
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port + (7 if mode == "check_causal" else 0)), os.path.join(ROOT, "tools", "ring_bench.py"), mode]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["n_gpus"] == n and res["ok"] and res["max_abs_vs_single_device"] < 6e-3, res
