"""probe: host-to-device bandwidth of 96 MB from pinned memory, split over 1..8 concurrent streams, and with a compute kernel running"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
n = 12_000_000
h = torch.empty(n, dtype=torch.float64).pin_memory(); h.uniform_(-3, 3)
d = torch.empty(n, dtype=torch.float64, device="cuda")
for ns in (1, 2, 4, 8, 16):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    per = n // ns
    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * per:(i + 1) * per].copy_(h[i * per:(i + 1) * per], non_blocking=True)
    go(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): go()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 10
    print("H2D 96 MB over %2d streams: %.3f ms = %.1f GB/s" % (ns, 1e3 * t, 0.096 / t), flush=True)
# many small chunks on one stream
for chunk_mb in (1.5, 3, 6, 12, 24):
    per = int(chunk_mb * 1e6 / 8); k = n // per
    s = torch.cuda.Stream()
    def go2():
        with torch.cuda.stream(s):
            for i in range(k): d[i * per:(i + 1) * per].copy_(h[i * per:(i + 1) * per], non_blocking=True)
    go2(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): go2()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 10
    print("H2D %d x %.1f MB on one stream: %.3f ms = %.1f GB/s" % (k, chunk_mb, 1e3 * t, k * per * 8e-9 / t), flush=True)
# with a busy GPU
x = torch.randn(8192, 8192, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def go3():
    with torch.cuda.stream(s1):
        for _ in range(6): y = x @ x
    with torch.cuda.stream(s2): d.copy_(h, non_blocking=True)
go3(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s1):
    for _ in range(20): y = x @ x
with torch.cuda.stream(s2):
    e0.record(); d.copy_(h, non_blocking=True); e1.record()
torch.cuda.synchronize()
print("H2D 96 MB while a GEMM stream is busy: %.3f ms = %.1f GB/s" % (e0.elapsed_time(e1), 0.096 / (1e-3 * e0.elapsed_time(e1))))
