"""Run-to-run reproducibility probe: n songs x secs seconds, `runs` analyses, reports which feature columns differ."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bliss_rs_b200 as B
from bliss_rs_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
secs = int(sys.argv[2]) if len(sys.argv) > 2 else 30
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
B.native.init(0)
pcm, offs, lens = synth.gen_corpus_flat(4242, list(range(n)), [22050 * secs] * n, device=dev)
outs = []
for _ in range(runs):
    out = torch.zeros((n, 23), device=dev)
    st = B.native.analyze_batch_device(pcm.data_ptr(), offs, lens, 2, out.data_ptr())
    torch.cuda.synchronize()
    outs.append(out.cpu())
for r in outs[1:]:
    d = (r != outs[0])
    print("differ:", int(d.sum()), "columns", d.any(0).nonzero().flatten().tolist(), "rows", d.any(1).nonzero().flatten().tolist()[:20])
print("tempo run0:", [round(float(v), 5) for v in outs[0][:8, 0]])
print("tempo run1:", [round(float(v), 5) for v in outs[-1][:8, 0]])
