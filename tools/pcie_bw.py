import torch, time
n = 64 << 20
hs = torch.empty(n, dtype=torch.uint8).pin_memory(); hd = torch.empty(n, dtype=torch.uint8).pin_memory()
ds = torch.empty(n, dtype=torch.uint8, device="cuda"); dd = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, reps=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): ds.copy_(hs, non_blocking=True)
        if down:
            with torch.cuda.stream(s2): hd.copy_(dd, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    return n / dt / 1e9
run(True, True, 3)
print("H2D alone %.1f GB/s" % run(True, False)); print("D2H alone %.1f GB/s" % run(False, True))
b = run(True, True); print("both: %.1f GB/s each, %.1f summed" % (b, 2 * b))
