"""PCIe ceiling of the box: H2D alone, D2H alone, both at once (pinned 1 GiB buffers, CUDA events)."""
import torch, time
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def up():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def down():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): up(); down()
a, b, c = t(up), t(down), t(both)
print("H2D alone %.1f GB/s   D2H alone %.1f GB/s   both at once: %.1f GB/s each, %.1f GB/s aggregate" % (n / a / 1e9, n / b / 1e9, n / c / 1e9, 2 * n / c / 1e9))
