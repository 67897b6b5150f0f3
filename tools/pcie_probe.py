#!/usr/bin/env python
"""PCIe probe at the copy sizes of the e2e leg (development tool): per 'frame' 4 H2D copies (2.76 + 3 x 1.84 MB, 16-bit
planes) and 2 D2H copies (2 x 3.69 MB), on separate streams, alone and together, with and without a concurrent kernel."""
import time

import torch

dev = torch.device("cuda", 0)
npx = 1280 * 720
h_in = [torch.empty(s, dtype=torch.uint8, pin_memory=True) for s in (3 * npx, 2 * npx, 2 * npx, 2 * npx)]
d_in = [torch.empty_like(t, device=dev) for t in h_in]
d_out = [torch.empty(4 * npx, dtype=torch.uint8, device=dev) for _ in range(2)]
h_out = [torch.empty(4 * npx, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
sa, sb, sc = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
big = torch.empty(256 << 20, dtype=torch.float32, device=dev)


def run(h2d, d2h, kernel, iters=400):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        if h2d:
            with torch.cuda.stream(sa):
                for h, d in zip(h_in, d_in):
                    d.copy_(h, non_blocking=True)
        if d2h:
            with torch.cuda.stream(sb):
                for h, d in zip(h_out, d_out):
                    h.copy_(d, non_blocking=True)
        if kernel:
            with torch.cuda.stream(sc):
                big.mul_(1.0001)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    bi = sum(t.numel() for t in h_in) * iters if h2d else 0
    bo = sum(t.numel() for t in h_out) * iters if d2h else 0
    print(f"h2d={h2d} d2h={d2h} kernel={kernel}: {1e6 * dt / iters:7.1f} us/frame  in {bi / dt / 1e9:5.1f} GB/s  out {bo / dt / 1e9:5.1f} GB/s", flush=True)


for args in ((1, 0, 0), (0, 1, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1), (0, 1, 1)):
    run(*args)

print("--- duplex vs copy size (one H2D + one D2H copy per iteration, equal sizes)")
for mb in (1, 2, 4, 8, 16, 64, 256):
    n = mb << 20
    hi, ho = torch.empty(n, dtype=torch.uint8, pin_memory=True), torch.empty(n, dtype=torch.uint8, pin_memory=True)
    di, do = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    iters = max(4, 2048 // mb)
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(sa):
                    di.copy_(hi, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(sb):
                    ho.copy_(do, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{mb:4d} MB {mode:5s}: {n * iters / dt / 1e9:5.1f} GB/s per direction", flush=True)
