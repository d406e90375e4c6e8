import torch, time
n = 1 << 28  # 1 GiB of float32
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.ones(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
for name, fn, gb in (("h2d", h2d, 4 * n / 1e9), ("d2h", d2h, 4 * n / 1e9), ("both", both, 8 * n / 1e9)):
    t(fn, 1); dt = t(fn)
    print(f"{name}: {dt*1e3:.1f} ms  {gb/dt:.1f} GB/s")
