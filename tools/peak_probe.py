"""Measure the int8 dense tensor peak the same way MEASURED_PEAKS.json measures bf16 (library GEMM, 8192^3,
best of 10, CUDA events) so int8 kernels have a measured roofline denominator.  Writes gpurun_out/peaks.json."""
import json
import os

import torch


def best_ms(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    N = 8192
    out = {"gpu": torch.cuda.get_device_name(0)}
    a = torch.randint(-127, 127, (N, N), dtype=torch.int8, device="cuda")
    b = torch.randint(-127, 127, (N, N), dtype=torch.int8, device="cuda").t()
    try:
        ms = best_ms(lambda: torch._int_mm(a, b))
        out["int8_tops_int_mm"] = 2 * N ** 3 / ms / 1e9
    except Exception as e:  # noqa: BLE001
        out["int8_error"] = repr(e)
    x = torch.randn(N, N, dtype=torch.bfloat16, device="cuda")
    y = torch.randn(N, N, dtype=torch.bfloat16, device="cuda")
    ms = best_ms(lambda: torch.matmul(x, y))
    out["bf16_tflops_matmul"] = 2 * N ** 3 / ms / 1e9
    n = 1 << 30
    s = torch.empty(n, dtype=torch.bfloat16, device="cuda"); d = torch.empty_like(s)
    ms = best_ms(lambda: d.copy_(s))
    out["hbm_copy_gbs"] = 2 * 2 * n / ms / 1e6
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/peaks.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
