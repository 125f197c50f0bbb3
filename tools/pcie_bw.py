#!/usr/bin/env python
"""Pinned host<->device copy rates on this box (16 MiB, the host-step payload), alone and both directions at once."""
import torch, time
n = 1 << 24
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for name, a, b in (("H2D", 1, 0), ("D2H", 0, 1), ("both", 1, 1)):
    run(a, b, 3); ms = run(a, b)
    print(f"{name:5s} 16 MiB: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s per direction")
