"""Host->device copy bandwidth of the box (pinned memory, one stream): the ceiling of the e2e metric.
Sweeps the copy size, and repeats the sweep while the analysis kernels run on another stream (do copies
slow down next to HBM-heavy kernels?)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

total = 4 << 30
h = torch.empty(total, dtype=torch.uint8, pin_memory=True)
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
copy_stream = torch.cuda.Stream()


def sweep(tag, busy=None):
    res = {}
    for mb in (16, 64, 256, 1024):
        n = mb << 20
        reps = total // n
        torch.cuda.synchronize()
        if busy:
            busy()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(copy_stream):
            e0.record()
            for r in range(reps):
                d[:n].copy_(h[r * n:(r + 1) * n], non_blocking=True)
            e1.record()
        torch.cuda.synchronize()
        res["%dMB" % mb] = round(total / (e0.elapsed_time(e1) / 1e3) / 1e9, 2)
    return res


out = {"bytes_per_sweep": total, "idle_gpu_gbs": sweep("idle")}
import bliss_rs_b200 as B  # noqa: E402
from bliss_rs_b200 import synth  # noqa: E402
nat = B.native
nat.init(0)
S = 256
pcm, offs, lens = synth.gen_corpus_flat(1, list(range(S)), [3969000] * S, device="cuda")
feats = torch.zeros((S, 23), device="cuda")
st = torch.cuda.current_stream().cuda_stream
nat.analyze_batch_device(pcm.data_ptr(), offs, lens, 2, feats.data_ptr(), st)
torch.cuda.synchronize()


def busy():
    for _ in range(6):  # ~120 ms of kernels queued on the default stream
        nat.analyze_batch_device(pcm.data_ptr(), offs, lens, 2, feats.data_ptr(), st)


out["under_analysis_kernels_gbs"] = sweep("busy", busy)
print(json.dumps(out))
