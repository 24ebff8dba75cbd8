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
    ap.add_argument("--cufft", action="store_true",
                    help="also time the on-box library comparator: torch.stft (cuFFT batched R2C) + abs on the same tracks")
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
    lib = None
    if args.cufft:
        # Library comparator (cuFFT for the STFT stage).  Same framing: PVocTempo's frame m is
        # x[256 m - 256 .. 256 m + 256) with zeros before the song, i.e. torch.stft(center=False) of the song behind
        # 256 zeros; hanningz window; magnitudes materialised.  Chunks of tracks bound the frame buffers torch
        # materialises (3 GB of windowed frames + 3 GB of complex spectra per 100 tracks).
        win = (0.5 * (1.0 - torch.cos(2.0 * torch.pi * torch.arange(512, device=dev, dtype=torch.float32) / 512.0)))
        chunk = 100

        def lib_pass():
            last = None
            for lo in range(0, R, chunk):
                x = pcm[lo * TRACK:min(lo + chunk, R) * TRACK].view(-1, TRACK)
                x = torch.nn.functional.pad(x, (256, 0))
                last = torch.stft(x, 512, hop_length=256, window=win, center=False, return_complex=True).abs()
            return last

        ref = lib_pass()  # warm-up (cuFFT plan) + a spot check against this library's magnitudes of the last chunk
        last_lo = ((R - 1) // chunk) * chunk
        ours = mags.view(R, n_t, 257)[last_lo:last_lo + ref.shape[0]]
        dev_err = float((ref[:, :, :n_t].transpose(1, 2) - ours).abs().max() / ours.abs().max())
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        lib_pass()
        c1.record()
        torch.cuda.synchronize()
        lib = {"impl": "torch.stft (cuFFT R2C, frames and complex spectra materialised) + abs", "tracks": R,
               "tracks_per_s": R / (c0.elapsed_time(c1) / 1e3), "max_diff_rel_to_max": dev_err}
    print(json.dumps({
        "library_comparator": lib,
        "bench": "stft512_hop256_mag", "kernel_variant_mask": int(os.environ.get("BLISS_B200_VARIANT", "0") or 0), "tracks": tracks, "resident_tracks": R, "passes": passes,
        "ms_total": ms, "tracks_per_s": tracks / (ms / 1e3), "read_gbs_algorithmic": rd, "write_gbs": wr,
        "hbm_peak_gbs": peak, "peak_source": src, "frac_read_of_peak": rd / peak,
        "frac_read_plus_write_of_peak": (rd + wr) / peak,
        "note": "input (%.1f GB) >> L2; each pass re-reads every sample from HBM" % (R * TRACK * 4 / 1e9)}))


if __name__ == "__main__":
    main()
