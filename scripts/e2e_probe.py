"""e2e experiment: bliss_b200_analyze_batch on pinned host buffers for several chunk sizes, with the
library's own stream trace (BLISS_B200_TRACE) -- run once per chunk size in a fresh process."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bliss_rs_b200 as B
from bliss_rs_b200 import synth
TRACK = 3969000
ES = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nat = B.native
nat.init(0)
base = [synth.gen_track(5, i, TRACK, "cuda") for i in range(16)]
host = torch.empty(ES * TRACK, dtype=torch.float32, pin_memory=True)
for i in range(ES):
    host[i * TRACK:(i + 1) * TRACK].copy_(base[i % 16])
torch.cuda.synchronize()
ptrs = (ctypes.c_void_p * ES)(*[host.data_ptr() + 4 * i * TRACK for i in range(ES)])
lens = (ctypes.c_uint64 * ES)(*([TRACK] * ES))
out = np.zeros((ES, 23), np.float32)
st = np.zeros(ES, np.int32)
nat.analyze_batch_ptrs(ptrs, lens, 2, out, st)
for rep in range(3):
    t0 = time.perf_counter()
    nat.analyze_batch_ptrs(ptrs, lens, 2, out, st)
    dt = time.perf_counter() - t0
    print("chunk_mb=%s songs=%d wall=%.1f ms -> %.0f songs/s, %.1f GB/s" % (os.environ.get("BLISS_B200_CHUNK_MB", "default"), ES, dt * 1e3, ES / dt, ES * TRACK * 4 / dt / 1e9), flush=True)
