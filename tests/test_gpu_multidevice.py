"""bliss_b200_init_devices: ONE process drives every B200 of the box through the host-buffer C ABI (the reference is
one process: worker threads + a channel, src/song/decoder.rs:282-331).  Needs >= 2 GPUs (skipped otherwise); run in a
child process because the context table of a process is bound at its first init."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, time
import numpy as np
sys.path.insert(0, %r)
import torch
import bliss_rs_b200 as B
from bliss_rs_b200 import synth
nat = B.native
n_gpu = torch.cuda.device_count()
assert n_gpu >= 2, n_gpu
rng = np.random.default_rng(7)
# ragged corpus: 30 s .. 4 min, two rejected songs
lens = [int(22050 * s) for s in rng.uniform(30, 240, 22)] + [100, 0, 8192, 22050 * 200]
songs = [synth.gen_track(99, i, max(n, 1), device="cuda:0").cpu().numpy()[:n] for i, n in enumerate(lens)]
nat.init(0)
st1, f1 = nat.analyze_batch(songs, 2)
d1 = nat.distance_matrix(f1[:20], f1[:20])
import ctypes
big = np.tile(f1[:20], (120, 1)).astype(np.float32)         # 2400 x 23: large enough for the row-block split
dm1 = nat.distance_matrix(big, big)
n = nat.init_devices(0)
assert n == n_gpu and nat.device_count() == n_gpu, (n, n_gpu)
st2, f2 = nat.analyze_batch(songs, 2)
assert np.array_equal(st1, st2), (st1, st2)
assert np.array_equal(f1, f2), np.abs(f1 - f2).max()
dm2 = nat.distance_matrix(big, big)
assert np.array_equal(dm1, dm2)
# s16 ingest through every device
s16 = [(x * 32767.0).round().astype(np.int16) for x in songs]
st3, f3 = nat.analyze_batch_s16(s16, 2)
nat_single = [nat.analyze_batch([ (s.astype(np.float32) / np.float32(32768.0)) ], 2)[1][0] if len(s) >= 8192 else None for s in s16[:4]]
for i, want in enumerate(nat_single):
    if want is not None:
        assert np.array_equal(f3[i], want), i
# throughput of the one call from pinned host memory (reported, not asserted)
L = 22050 * 60
host = torch.empty(64 * L, dtype=torch.float32).pin_memory()
host.copy_(torch.from_numpy(np.tile(songs[0][:L] if len(songs[0]) >= L else np.resize(songs[0], L), 64)))
ptrs = (ctypes.c_void_p * 64)(*[host.data_ptr() + 4 * i * L for i in range(64)])
hl = (ctypes.c_uint64 * 64)(*([L] * 64))
out = np.zeros((64, 23), np.float32); stt = np.zeros(64, np.int32)
nat.analyze_batch_ptrs(ptrs, hl, 2, out, stt)
t0 = time.perf_counter(); nat.analyze_batch_ptrs(ptrs, hl, 2, out, stt); dt = time.perf_counter() - t0
print("MULTI_OK devices=%%d  64 x 1-min songs through one call: %%.1f songs/s (%%.2f GB/s H2D)" %% (n, 64 / dt, 64 * L * 4 / dt / 1e9))
''' % ROOT


@pytest.mark.gpu
def test_one_process_drives_every_device():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (bliss_b200_init_devices shards one call's songs over all devices)")
    out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(out.stdout[-2000:])
    assert out.returncode == 0 and "MULTI_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
