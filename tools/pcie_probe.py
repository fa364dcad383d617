"""Host<->device copy bandwidth of this box (pinned memory), alone and full duplex: the floor for bench.py's e2e."""
import json

import torch

n = 1 << 29   # 512 MiB
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def h2d():
    with torch.cuda.stream(s1):
        s1.wait_stream(torch.cuda.current_stream()); d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        s2.wait_stream(torch.cuda.current_stream()); h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


res = {"h2d_GBs": n / timed(h2d) / 1e6, "d2h_GBs": n / timed(d2h) / 1e6}
t = timed(both)
res["duplex_each_GBs"] = n / t / 1e6
print(json.dumps(res))
