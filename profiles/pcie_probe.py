"""Pinned host<->device copy bandwidth on this box: H2D alone, D2H alone, both at once (two streams).
The e2e figure of bench.py is bounded by these.  usage: python profiles/pcie_probe.py [GiB]"""
import json, sys, time
import torch
gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
n = int(gib * (1 << 30)) // 8
h_in = torch.empty(n, dtype=torch.float64, pin_memory=True); h_in.fill_(1.0)
h_out = torch.empty(n, dtype=torch.float64, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps

def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
def chunked(k=24):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)
gb = n * 8 / 1e9
res = {"GiB_each": gib, "h2d_GBps": gb / timed(h2d), "d2h_GBps": gb / timed(d2h)}
t = timed(both); res["both_GBps_each"] = gb / t; res["both_aggregate_GBps"] = 2 * gb / t
t = timed(chunked); res["both_chunked24_aggregate_GBps"] = 2 * gb / t
print(json.dumps(res))
