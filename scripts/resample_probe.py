#!/usr/bin/env python3
"""Small workload over every resampler kernel and the fused feed, for compute-sanitizer (scripts/gpu_r02_s.sh)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bliss_rs_b200 as B  # noqa: E402

nat = B.native
nat.init(0)
rng = np.random.default_rng(0)
for variant in (0, 524288):
    nat.set_variant(variant)
    for rate, n in ((44100, 30001), (88200, 50003), (48000, 40007), (96000, 30000), (32000, 20000), (8000, 5000), (11025, 7000),
                    (192000, 60000), (22051, 3000), (48000, 5), (44100, 3)):
        y = nat.resample(rng.standard_normal(n).astype(np.float32), rate)
        assert y.size == nat.resampled_len(n, rate)
nat.set_variant(0)
songs = [rng.integers(-20000, 20000, (n, 2), dtype=np.int16) for n in (90000, 40001, 17000, 0, 65536)]
for rate in (44100, 48000):
    st, f = nat.analyze_batch_pcm(songs, rate, 2)
    print(rate, list(st), float(np.abs(f).max()))
print("probe ok")
