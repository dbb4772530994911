"""Host <-> device copy bandwidth of this box from pinned memory (what bounds the e2e path's raw upload)."""
import sys
import time

import torch

dev = torch.device("cuda", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
torch.cuda.set_device(dev)
for mb in (16, 128, 1024):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"{name} {mb:5d} MB pinned: {n / dt / 1e9:6.1f} GB/s")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"both {mb:5d} MB pinned: {n / dt / 1e9:6.1f} GB/s each direction")
p = torch.empty(256 << 20, dtype=torch.uint8)
d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
d.copy_(p); torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(p); torch.cuda.synchronize()
print(f"h2d 256 MB pageable: {(256 << 20) / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
