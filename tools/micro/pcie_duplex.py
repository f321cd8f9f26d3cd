"""Does an H2D copy on one stream overlap a D2H copy on another (PCIe full duplex)?  Development probe for the e2e pipeline of bench.py."""
import torch

n_up, n_dn = 640_000_000, 1_316_000_000
h_up = torch.empty(n_up, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(n_dn, dtype=torch.uint8).pin_memory()
d_up = torch.empty(n_up, dtype=torch.uint8, device="cuda")
d_dn = torch.empty(n_dn, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, dn, dn_first=False):
    torch.cuda.synchronize()
    e0, e1, j = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event()
    e0.record(s1)
    s2.wait_event(e0)
    if dn and dn_first:
        with torch.cuda.stream(s2):
            h_dn.copy_(d_dn, non_blocking=True)
    if up:
        with torch.cuda.stream(s1):
            d_up.copy_(h_up, non_blocking=True)
    if dn and not dn_first:
        with torch.cuda.stream(s2):
            h_dn.copy_(d_dn, non_blocking=True)
    j.record(s2)
    s1.wait_event(j)
    e1.record(s1)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


for _ in range(2):
    run(True, True)
a, b, c = min(run(True, False) for _ in range(3)), min(run(False, True) for _ in range(3)), min(run(True, True) for _ in range(3))
print(f"H2D {n_up / 1e6:.0f} MB alone {a:.2f} ms ({n_up / a / 1e6:.1f} GB/s); D2H {n_dn / 1e6:.0f} MB alone {b:.2f} ms ({n_dn / b / 1e6:.1f} GB/s); "
      f"both at once {c:.2f} ms (sum {a + b:.2f}, max {max(a, b):.2f})")
d = min(run(True, True, dn_first=True) for _ in range(3))
print(f"D2H issued first, then H2D: {d:.2f} ms")
