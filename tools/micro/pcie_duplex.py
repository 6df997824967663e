"""PCIe ceiling of the box: pinned host <-> device copies one way, the other way, and both at the same time (the end-to-end leg's limit)."""
import time, torch
n = 1 << 30
h_up, h_dn = torch.empty(n, dtype=torch.uint8, pin_memory=True), torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_up, d_dn = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, dn, reps=6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1): d_up.copy_(h_up, non_blocking=True)
        if dn:
            with torch.cuda.stream(s2): h_dn.copy_(d_dn, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * n / dt / 1e9
run(True, True, 2)
print("H2D alone %.1f GB/s, D2H alone %.1f GB/s, both at once %.1f GB/s each way" % (run(True, False), run(False, True), run(True, True)))
