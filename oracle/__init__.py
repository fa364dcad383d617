"""CPU oracle for the three attention paths of selau642/QuantizedAttention.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this package.  The product
package `quantizedattention_b200` never imports it and has no CPU fallback.

Contents (each function cites the reference file:line it restates):
  baseline.py  -- `baseline_pytorch_attention` x3 (fp32 PyTorch math = oracle of intent and the
                  CPU baseline of record, attention_int8.py:453-481, attention_bf16.py:450-478,
                  attention_jvp.py:197-215)
  int8_ref.py  -- block quantisation, int8 forward, int8 backward (`literal` = bug-for-bug,
                  `contract` = SURVEY.md 8-LEDGER fixes)
  bf16_ref.py  -- bias-corrected bf16 forward and the fp32 "Algorithm 4" backward
  jvp_ref.py   -- forward-mode JVP attention
  _helion_standin/ -- eager stand-in for the un-installable `helion` package so that the
                  UNMODIFIED /root/reference/*.py can be executed on CPU to pin the above
  make_golden.py -- runs the real reference under the stand-in and writes tests/golden/*.pt

Parity pinning: the reference ships NO golden vectors, KATs or assertions.
The oracle is pinned instead against outputs of the reference itself: `make_golden.py`
executes the unmodified reference source (eager semantics via the stand-in) and the committed
fixtures in tests/golden/ are compared bit-for-bit (`torch.equal`) with the `literal`
restatements in tests/test_oracle_golden.py.  The Helion->Triton code generator's own
intermediate roundings are NOT pinned (Helion cannot be installed).
"""
