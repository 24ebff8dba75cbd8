#!/usr/bin/env python3
"""BASELINE.json configs[4]: a mixed-duration corpus (30 s - 10 min, Zipf) end to end through ONE
bliss_b200_analyze_batch call of ONE process over every GPU of the box (bliss_b200_init_devices: the library deals the
songs longest-first over the devices), then playlist-from-seed (closest_to_songs, src/playlist.rs:256-270) against the
oracle's ordering on a subset.

  python bench.py --config 5 --gpus 8            (or: python bench_config5.py --gpus 8 --songs 20000)

Corpus (SURVEY.md section 8d "Configs 4-5"): duration_i = 30 s x k_i, k_i ~ Zipf(alpha) (numpy Generator.zipf, seed
stated), clipped to [30 s, 10 min]; N_i = round(22050 x duration_i).  The PCM of song i is a slice (random start, 16-byte
aligned) of one of `--pool` distinct synthetic 10-minute tracks held in pinned host memory: songs share host bytes but no
two are the same signal.  Timed (wall clock around the call, host buffers in, 23 floats per song out): songs/s and
audio-seconds/s.  Reported beside it: the longest-first load balance over the devices (max / mean samples per device),
feature parity of the subset against the oracle (max / median error per feature, tempo flips) and the order of
closest_to_songs([seed], subset) against the oracle's order on ITS features (first-k exact, Kendall tau; ties fall to
the input index on both sides: the reference sort is stable, src/playlist.rs:267-268).
Under torchrun only rank 0 works (one process is the point); the other ranks exit.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
SR = 22050
POOL_SAMPLES = 10 * 60 * SR


def zipf_lengths(n, alpha, seed):
    rng = np.random.default_rng(seed)
    k = rng.zipf(alpha, size=n).astype(np.float64)
    dur = np.clip(30.0 * k, 30.0, 600.0)
    return np.round(dur * SR).astype(np.int64), rng


def lpt_loads(lengths, n_dev):
    """the library's deal (api.cu shard_lpt): longest first onto the least loaded device"""
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(n_dev, np.int64)
    cnt = np.zeros(n_dev, np.int64)
    for i in order:
        d = int(np.argmin(load))
        load[d] += int(lengths[i]) + 4096
        cnt[d] += 1
    return load, cnt


def kendall_tau(a, b):
    from scipy.stats import kendalltau
    return float(kendalltau(a, b).statistic)


def main(argv=None, emit=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0, help="devices to use (0 = all visible)")
    ap.add_argument("--songs", type=int, default=20000)
    ap.add_argument("--alpha", type=float, default=1.5)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--pool", type=int, default=24, help="distinct 10-minute tracks the songs are cut from")
    ap.add_argument("--subset", type=int, default=2048, help="songs also analysed by the CPU oracle (parity + playlist order)")
    ap.add_argument("--first-k", type=int, default=50)
    args = ap.parse_args(argv)
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    import torch
    import bliss_rs_b200 as B
    from bliss_rs_b200 import synth
    from oracle import oracle as O
    nat = B.native
    n_dev = nat.init_devices(args.gpus)
    lengths, rng = zipf_lengths(args.songs, args.alpha, args.seed)
    n = len(lengths)
    # the pool of distinct tracks: generated on GPU 0, kept in pinned host memory
    pool = torch.empty((args.pool, POOL_SAMPLES), dtype=torch.float32, pin_memory=True)
    for i in range(args.pool):
        pool[i].copy_(synth.gen_track(args.seed, i, POOL_SAMPLES, device="cuda:0"))
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    starts = (rng.integers(0, POOL_SAMPLES - lengths + 1) // 4) * 4
    which = np.arange(n) % args.pool
    base = pool.data_ptr()
    ptrs = (ctypes.c_void_p * n)(*[base + 4 * (int(which[i]) * POOL_SAMPLES + int(starts[i])) for i in range(n)])
    lens = (ctypes.c_uint64 * n)(*[int(v) for v in lengths])
    out = np.zeros((n, 23), np.float32)
    status = np.zeros(n, np.int32)
    # warm-up on a slice of the corpus (allocations on every device), then the timed call over all of it
    w = min(n, 64 * n_dev)
    nat.analyze_batch_ptrs((ctypes.c_void_p * w)(*ptrs[:w]), (ctypes.c_uint64 * w)(*lens[:w]), 2, out[:w], status[:w])
    t0 = time.perf_counter()
    nat.analyze_batch_ptrs(ptrs, lens, 2, out, status)
    dt = time.perf_counter() - t0
    audio_s = float(lengths.sum()) / SR
    load, cnt = lpt_loads(lengths, n_dev)
    # ---- the subset through the oracle: parity and playlist order ------------------------------------------------------
    m = min(args.subset, n)
    host = pool.numpy()
    songs = [host[which[i], starts[i]:starts[i] + lengths[i]] for i in range(m)]
    cores = os.cpu_count() or 1
    tc = time.perf_counter()
    ost, ofe = O.analyze_batch(songs, 2, n_threads=cores)
    dtc = time.perf_counter() - tc
    gf = out[:m]
    err = np.abs(gf - ofe)
    tol = 1e-4 * np.maximum(1.0, np.abs(ofe))
    order_gpu, _ = nat.closest_to_songs(gf[:1], gf)
    order_ref, _ = O.closest_to_songs(ofe[:1], ofe)
    pos_gpu, pos_ref = np.empty(m, np.int64), np.empty(m, np.int64)
    pos_gpu[order_gpu.astype(np.int64)] = np.arange(m)
    pos_ref[order_ref.astype(np.int64)] = np.arange(m)
    k = min(args.first_k, m)
    # and the whole corpus once through the device ordering (timed for the record)
    tp = time.perf_counter()
    order_all, _ = nat.closest_to_songs(out[:1], out)
    dtp = time.perf_counter() - tp
    line = {
        "bench": "configs[4]: mixed-duration corpus, one multi-device call end to end + playlist from seed",
        "n_gpus": n_dev, "songs": n, "zipf_alpha": args.alpha, "seed": args.seed,
        "duration_s": {"min": float(lengths.min()) / SR, "median": float(np.median(lengths)) / SR,
                       "mean": float(lengths.mean()) / SR, "max": float(lengths.max()) / SR},
        "host_bytes": int(lengths.sum()) * 4, "wall_s": dt,
        "e2e_songs_per_s": n / dt, "e2e_audio_seconds_per_s": audio_s / dt, "h2d_gbs": float(lengths.sum()) * 4 / 1e9 / dt,
        "all_ok": bool((status == 0).all()),
        "lpt": {"samples_per_device": [int(v) for v in load], "songs_per_device": [int(v) for v in cnt],
                "imbalance_max_over_mean": float(load.max() / load.mean())},
        "parity_subset": {"songs": m, "oracle_threads": cores, "oracle_wall_s": dtc, "oracle_songs_per_s": m / dtc,
                          "max_abs_err": float(err.max()), "within_1e-4": bool((err <= tol).all()),
                          "per_feature_max_abs_err": [float(v) for v in err.max(0)],
                          "per_feature_median_abs_err": [float(v) for v in np.median(err, 0)],
                          "tempo_flips": int((err[:, 0] > 1e-3).sum())},
        "playlist_from_seed": {"candidates": m, "first_k": k,
                               "first_k_exact": bool(np.array_equal(order_gpu[:k], order_ref[:k])),
                               "first_mismatch_rank": int(np.argmax(order_gpu != order_ref)) if (order_gpu != order_ref).any() else None,
                               "kendall_tau": kendall_tau(pos_gpu, pos_ref),
                               "whole_corpus_order_ms": dtp * 1e3, "whole_corpus_first": [int(v) for v in order_all[:5]]},
        "note": "wall clock around ONE bliss_b200_analyze_batch call (host pointers in, features out); the songs are "
                "slices of %d pinned 10-minute tracks" % args.pool,
    }
    (emit or (lambda d: print(json.dumps(d), flush=True)))(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
