"""Host->device copy bandwidth of the box (pinned memory, one stream): the ceiling of the e2e metric."""
import json
import time

import torch

n = 1 << 30  # 1 GiB
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 8
for _ in range(reps):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"h2d_pinned_gbs": reps * n / dt / 1e9, "bytes": n, "reps": reps}))
