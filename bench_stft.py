#!/usr/bin/env python3
"""STFT-only micro-benchmark (BASELINE.json configs[2]): 512-point hanningz window, hop 256
(= PVocTempo framing, src/aubio.rs:338-425), 257 magnitudes per frame materialised in HBM, over
10 000 synthetic 3-min tracks (a resident subset is looped; stated in the output).

Prints one JSON line: tracks/s, algorithmic read GB/s (4 N bytes per track: every sample once),
read+write GB/s, fraction of the measured HBM peak.  Device-timed with CUDA events.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
TRACK = 3 * 60 * 22050


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=10000)
    ap.add_argument("--resident", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=1)
    args = ap.parse_args()
    import torch
    import bliss_rs_b200 as B
    from bliss_rs_b200 import synth
    nat = B.native
    nat.init(0)
    dev = torch.device("cuda", 0)
    R = min(args.resident, args.tracks)
    distinct = 32
    base = [synth.gen_track(7, i, TRACK, dev) for i in range(distinct)]
    pcm = torch.empty(R * TRACK, dtype=torch.float32, device=dev)
    for i in range(R):
        pcm[i * TRACK:(i + 1) * TRACK].copy_(base[i % distinct] * (0.5 + 0.5 * ((i * 7919) % 97) / 97.0))
    del base
    n_t = (TRACK - 512) // 256 + 1
    mags = torch.empty((R * n_t, 257), dtype=torch.float32, device=dev)
    offs = [i * TRACK for i in range(R)]
    lens = [TRACK] * R
    stream = torch.cuda.current_stream().cuda_stream
    passes = -(-args.tracks // R)
    for _ in range(args.warmup):
        nat.stft512_mag_device(pcm.data_ptr(), offs, lens, mags.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(passes):
        nat.stft512_mag_device(pcm.data_ptr(), offs, lens, mags.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tracks = passes * R
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        src = "measured"
    except Exception:
        peak, src = 6650.0, "fallback"
    rd = tracks * TRACK * 4 / 1e9 / (ms / 1e3)
    wr = tracks * n_t * 257 * 4 / 1e9 / (ms / 1e3)
    print(json.dumps({
        "bench": "stft512_hop256_mag", "kernel_variant_mask": int(os.environ.get("BLISS_B200_VARIANT", "0") or 0), "tracks": tracks, "resident_tracks": R, "passes": passes,
        "ms_total": ms, "tracks_per_s": tracks / (ms / 1e3), "read_gbs_algorithmic": rd, "write_gbs": wr,
        "hbm_peak_gbs": peak, "peak_source": src, "frac_read_of_peak": rd / peak,
        "frac_read_plus_write_of_peak": (rd + wr) / peak,
        "note": "input (%.1f GB) >> L2; each pass re-reads every sample from HBM" % (R * TRACK * 4 / 1e9)}))


if __name__ == "__main__":
    main()
