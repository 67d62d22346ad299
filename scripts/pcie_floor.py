"""PCIe floor of the host-streamed ray stream on this box: pinned H2D, D2H and both at once (two streams), 174 MB each."""
import time, torch
n = 2073600 * 84
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(f, reps=10):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
a, b, c = run(h2d), run(d2h), run(both)
print(f"PCIe floor for {n/1e6:.1f} MB: H2D {a:.3f} ms ({n/a/1e6:.1f} GB/s), D2H {b:.3f} ms ({n/b/1e6:.1f} GB/s), both at once {c:.3f} ms")
