#!/usr/bin/env python3
"""Diagnostic: which config-5 songs miss the 1e-4 feature bar and why (per-frame roll-off against the oracle, round-1
and round-2 pvoc kernels)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bliss_rs_b200 as B
from bliss_rs_b200 import synth
from oracle import oracle as O
import bench_config5 as C5
nat = B.native
nat.init(0)
n, m = 3000, int(sys.argv[1]) if len(sys.argv) > 1 else 512
lengths, rng = C5.zipf_lengths(n, 1.5, 20261017)
pool = [synth.gen_track(20261017, i, C5.POOL_SAMPLES, device="cuda:0").cpu().numpy() for i in range(24)]
starts = (rng.integers(0, C5.POOL_SAMPLES - lengths + 1) // 4) * 4
songs = [pool[i % 24][starts[i]:starts[i] + lengths[i]] for i in range(m)]
_, f2 = nat.analyze_batch(songs, 2)
nat.set_variant(16384)
_, f1 = nat.analyze_batch(songs, 2)
nat.set_variant(0)
_, ofe = O.analyze_batch(songs, 2, n_threads=os.cpu_count())
for name, f in (("v2", f2), ("v1", f1)):
    err = np.abs(f - ofe)
    worst = np.argsort(-err.max(1))[:5]
    print(name, "max err per feature[:10]", np.array2string(err.max(0)[:10], precision=2))
    print(name, "worst songs", [(int(i), int(lengths[i]), float(err[i].max()), int(err[i].argmax())) for i in worst])
i = int(np.argsort(-np.abs(f2 - ofe).max(1))[0])
x = songs[i]
c, r, fl = O.timbral_frames(x)
for name, mask in (("v2", 0), ("v1", 16384)):
    nat.set_variant(mask)
    _, _, t = nat.analyze_taps(x, 2)
    d = (t["rolloff"] - r) / (22050 / 512)
    print(name, "song", i, "frames", r.size, "rolloff mismatching frames %.4f" % np.mean(d != 0), "max |bins| %d" % np.abs(d).max(),
          "mean signed bins %.4f" % d.mean(), "hist of |d| (0,1,2,3-5,6+):",
          [int((np.abs(d) == 0).sum()), int((np.abs(d) == 1).sum()), int((np.abs(d) == 2).sum()), int(((np.abs(d) >= 3) & (np.abs(d) <= 5)).sum()), int((np.abs(d) >= 6).sum())])
    ce = np.abs(t["centroid"] - c) / np.maximum(1, np.abs(c))
    print(name, "centroid max rel err %.2e, flatness max err/bar %.2f" % (ce.max(), (np.abs(t["flatness"] - fl) / (2e-4 * np.abs(fl) + 1e-5)).max()))
    bad = np.nonzero(np.abs(d) >= 3)[0][:5]
    print(name, "examples (frame, gpu bin, oracle bin):", [(int(k), float(t["rolloff"][k] / (22050 / 512)), float(r[k] / (22050 / 512))) for k in bad])
nat.set_variant(0)
