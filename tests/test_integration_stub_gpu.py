"""The reference-side ctypes stub printed in INTEGRATION.md is executed as written (against the in-tree libqattn.so) and must
give the same tensors as the package's own wrappers: the document is the binding a maintainer would paste."""
import os
import re

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_namespace():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"## ctypes stub.*?```python\n(.*?)```", text, re.S).group(1)
    code = code.replace('ctypes.CDLL("libqattn.so")', 'ctypes.CDLL(%r)' % os.path.join(ROOT, "quantizedattention_b200", "libqattn.so"))
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    ns["_L"].qa_last_error.restype = __import__("ctypes").c_char_p
    ns["_L"].qa_k_mean_workspace_bytes.restype = __import__("ctypes").c_size_t
    return ns


def test_stub_quant_block_and_fp4_forward_match_the_package():
    from quantizedattention_b200 import attention_fp4, ops
    ns = _stub_namespace()
    g = torch.Generator().manual_seed(3)
    q, k, v = [torch.randn(1, 2, 256, 128, generator=g).half().cuda() for _ in range(3)]
    qi, sq = ns["quant_block"](q, 128)
    qi2, sq2 = ops.quant_block(q, 128)
    assert torch.equal(qi, qi2.view_as(qi)) and torch.equal(sq, sq2)
    O = ns["sage_fp4"](q, k, v)
    torch.cuda.synchronize()
    assert torch.equal(O, attention_fp4.sage_attention_3_fp4(q, k, v))
