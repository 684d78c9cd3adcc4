"""H2D bandwidth of pinned host memory on this box: alone, split over two streams, and while the GPU is busy (debug aid for e2e)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

node = bench.pin_to_gpu_numa_node(0) if "--pin" in sys.argv else None
n = 796 * 2**20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda:0")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=6):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9


def one():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)


def two():
    half = n // 2
    with torch.cuda.stream(s1):
        d[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2):
        d[half:].copy_(h[half:], non_blocking=True)


def chunks(k):
    def f():
        c = n // k
        with torch.cuda.stream(s1):
            for i in range(k):
                d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
    return f


print("numa node", node, "one stream %.1f GB/s" % timed(one), "two streams %.1f GB/s" % timed(two), "2 chunks %.1f" % timed(chunks(2)), "128 chunks %.1f" % timed(chunks(128)))
# with the GPU busy: HBM-heavy elementwise work on another stream
a = torch.empty(2**30, dtype=torch.uint8, device="cuda:0")
b = torch.empty_like(a)
s3 = torch.cuda.Stream()
stop = [False]


def busy_one():
    with torch.cuda.stream(s3):
        for _ in range(40):
            b.copy_(a)
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)


print("while the GPU copies 80 GB inside HBM: %.1f GB/s" % timed(busy_one, reps=3))
x = torch.randn(8192, 8192, device="cuda:0", dtype=torch.bfloat16)


def busy_mm():
    with torch.cuda.stream(s3):
        for _ in range(60):
            torch.mm(x, x)
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)


print("while the SMs run bf16 GEMMs: %.1f GB/s" % timed(busy_mm, reps=3))
