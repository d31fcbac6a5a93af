"""PCIe copy bandwidth of the box: H2D alone, D2H alone, both directions at once (pinned host memory, 134 MB = one C3
volume per copy).  The e2e numbers of bench.py are bounded by these, not by the kernels."""
import json
import sys
import time

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 134217728
reps = 10
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_a.copy_(h_a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for _ in range(2):
    run(True, True)
t_h2d, t_d2h, t_both = run(True, False), run(False, True), run(True, True)
print(json.dumps({"bytes": n, "h2d_GBs": n / t_h2d / 1e9, "d2h_GBs": n / t_d2h / 1e9,
                  "both_each_direction_GBs": n / t_both / 1e9, "ms_h2d": t_h2d * 1e3, "ms_d2h": t_d2h * 1e3,
                  "ms_both": t_both * 1e3}))
