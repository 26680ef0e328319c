import torch, time
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    run(a, b, 2); t = run(a, b)
    print(f"{name}: {t*1e3:.2f} ms per 256 MiB each -> {(a + b) * n / t / 1e9:.1f} GB/s aggregate")
