#!/usr/bin/env python
"""Host <-> device copy bandwidth of the box (pinned memory), each direction alone and both at once."""
import torch, time
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return n * reps / dt / 1e9
run(True, True, 2)
print("H2D alone %.1f GB/s" % run(True, False)); print("D2H alone %.1f GB/s" % run(False, True))
print("both at once: %.1f GB/s each way" % run(True, True))
for mb in (1, 4, 14):
    m = mb << 20
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): h2[:m].copy_(d2[:m], non_blocking=True); torch.cuda.synchronize()
    print("D2H %d MB, synchronous: %.0f us per copy" % (mb, (time.perf_counter() - t0) / 50 * 1e6))
