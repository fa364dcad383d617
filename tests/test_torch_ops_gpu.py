"""torch.library registration: the fused paths as opaque operators with fake kernels and backward
formulas - identical numbers to the drop-in modules, traceable by torch.compile(fullgraph=True) (aot_eager backend:
no code generation involved) and clean under torch.library.opcheck."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(shape, seed, v_dtype=torch.float16):
    g = torch.Generator().manual_seed(seed)
    q, k, v, dO = [torch.randn(shape, generator=g) for _ in range(4)]
    return q.half().cuda(), k.half().cuda(), v.to(v_dtype).cuda(), dO.cuda()


def test_sage_op_matches_module_and_compiles():
    from quantizedattention_b200 import attention_int8 as A
    from quantizedattention_b200 import torch_ops as T
    q, k, v, dO = _inputs((1, 2, 256, 128), 5)
    ref = [t.clone().requires_grad_() for t in (q, k, v)]
    Or = A.sage_attention_3_int8(*ref)
    Or.backward(dO.half())

    def f(q, k, v):
        return T.sage_attention_3_int8_op(q * 1.0, k, v) * 2.0       # surrounded by ordinary ops

    cf = torch.compile(f, backend="aot_eager", fullgraph=True)        # fullgraph: a graph break would raise
    got = [t.clone().requires_grad_() for t in (q, k, v)]
    O = cf(*got)
    O.backward(dO.half() / 2.0)
    torch.cuda.synchronize()
    assert torch.equal(O / 2.0, Or)
    assert torch.equal(got[1].grad, ref[1].grad) and torch.equal(got[2].grad, ref[2].grad)
    assert (got[0].grad.float() - ref[0].grad.float()).abs().max() <= 2e-3 * ref[0].grad.abs().max() + 1e-6   # fp32 atomics order


def test_flash_bf16_op_matches_module_and_compiles():
    from quantizedattention_b200 import attention_bf16 as B
    from quantizedattention_b200 import torch_ops as T
    q, k, v, dO = _inputs((1, 2, 256, 64), 6, torch.bfloat16)
    for causal in (False, True):
        ref = [t.clone().requires_grad_() for t in (q, k, v)]
        Or = B.flash_atten_2_bf16(*ref, causal)
        Or.backward(dO)
        cf = torch.compile(lambda q, k, v: T.flash_atten_2_bf16_op(q, k, v, causal), backend="aot_eager", fullgraph=True)
        got = [t.clone().requires_grad_() for t in (q, k, v)]
        O = cf(*got)
        O.backward(dO)
        torch.cuda.synchronize()
        assert torch.equal(O, Or)
        for a, b in zip(got, ref):
            assert a.grad.dtype == b.grad.dtype
            assert (a.grad.float() - b.grad.float()).abs().max() <= 2e-3 * b.grad.float().abs().max() + 1e-6


def test_opcheck_schemas_and_fake_kernels():
    from quantizedattention_b200 import torch_ops as T
    q, k, v, dO = _inputs((1, 1, 128, 64), 7)
    tests = ("test_schema", "test_faketensor")
    torch.library.opcheck(T.sage_int8_fwd, (q, k, v, 128, 128, False), test_utils=tests)
    out = T.sage_int8_fwd(q, k, v, 128, 128, False)
    torch.library.opcheck(T.sage_int8_bwd, (dO.half(), *out, 128, 128, False), test_utils=tests)
    vb = v.to(torch.bfloat16)
    torch.library.opcheck(T.flash_bf16_fwd, (q, k, vb, True), test_utils=tests)
    O, lse = T.flash_bf16_fwd(q, k, vb, True)
    torch.library.opcheck(T.flash_bf16_bwd, (q, k, vb, O, lse, True, dO), test_utils=tests)


def test_lowp_ops_and_modules_match_and_compile():
    """fp8 / NVFP4 forwards as custom ops (fullgraph-compilable, opcheck-clean) and through the layout adapters."""
    from quantizedattention_b200 import attention_fp4, attention_fp8, modules
    from quantizedattention_b200 import torch_ops as T
    q, k, v, _ = _inputs((1, 2, 256, 128), 11)
    for op, ref in ((T.sage_attention_3_fp4_op, attention_fp4.sage_attention_3_fp4), (T.sage_attention_3_fp8_op, attention_fp8.sage_attention_3_fp8)):
        cf = torch.compile(lambda q, k, v: op(q * 1.0, k, v) * 2.0, backend="aot_eager", fullgraph=True)
        assert torch.equal(cf(q, k, v) / 2.0, ref(q, k, v))
    torch.library.opcheck(T.sage_fp4_fwd, (q, k, v, True), test_utils=("test_schema", "test_faketensor"))
    assert torch.equal(T.sage_attention_3_fp4_op(q, k, v, causal=True), attention_fp4.sage_attention_3_fp4(q, k, v, causal=True))
    assert torch.equal(modules.SageAttention3LowPrecision("fp4", layout="bhsd", causal=True)(q, k, v), attention_fp4.sage_attention_3_fp4(q, k, v, causal=True))
    torch.library.opcheck(T.sage_fp8_fwd, (q, k, v), test_utils=("test_schema", "test_faketensor"))
    for prec, ref in (("fp4", attention_fp4.sage_attention_3_fp4), ("fp8", attention_fp8.sage_attention_3_fp8)):
        m = modules.SageAttention3LowPrecision(prec, layout="bshd")
        got = m(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        assert got.shape == (1, 256, 2, 128) and torch.equal(got.transpose(1, 2), ref(q, k, v))
    with pytest.raises(ValueError):
        modules.sage_attention_lowp(q, k, v, precision="int3")
