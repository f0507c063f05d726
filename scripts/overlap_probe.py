"""Does an H2D copy on one stream overlap kernels on another on this box?"""
import time, torch
n = 128 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
x = torch.randn(64 << 20, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def kern():
    y = x
    for _ in range(12):
        y = y * 1.0001 + 0.5
    return y
for mode in ("copy", "kernel", "both"):
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if mode in ("copy", "both"):
            with torch.cuda.stream(s1):
                d.copy_(h, non_blocking=True)
        if mode in ("kernel", "both"):
            with torch.cuda.stream(s2):
                kern()
        torch.cuda.synchronize()
        t = (time.perf_counter() - t0) * 1e3
    print(mode, "%.3f ms" % t)
