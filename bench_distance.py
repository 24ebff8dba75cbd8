#!/usr/bin/env python3
"""All-pairs weighted distance (BASELINE.json configs[3] shape): n feature rows of 23 floats, this
rank's row block of the n x n matrix with the v2 metric (src/lib.rs:209-234, src/playlist.rs:140-142).
One JSON line: pairs/s, GB/s written vs the measured HBM peak.  Single GPU: the whole matrix in chunks
of rows (40 GB for n = 100 000)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--block-rows", type=int, default=12500)
    args = ap.parse_args()
    import torch
    import bliss_rs_b200 as B
    nat = B.native
    nat.init(0)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    feats = (torch.rand((args.n, 23), device=dev, generator=g) * 2 - 1).contiguous()
    w = nat.feature_weights(2)
    out = torch.empty((args.block_rows, args.n), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def block(lo):
        rows = feats[lo:lo + args.block_rows]
        nat.distance_matrix_device(rows.data_ptr(), rows.shape[0], feats.data_ptr(), args.n, 23, out.data_ptr(),
                                   nat.METRIC_MAHALANOBIS, w, stream)
        return rows.shape[0]

    block(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = 0
    for lo in range(0, args.n, args.block_rows):
        done += block(lo)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    pairs = done * args.n
    gbs = pairs * 4 / 1e9 / (ms / 1e3)
    # spot-check one entry against the direct formula
    i, j = 17, args.n - 3
    d = (feats[i] - feats[j]).double()
    ref = float(torch.sqrt((d * torch.tensor(w, device=dev).diagonal().double() * d).sum()))
    block(0)
    torch.cuda.synchronize()
    print(json.dumps({"bench": "all_pairs_distance_v2", "n": args.n, "pairs": pairs, "ms": ms, "pairs_per_s": pairs / (ms / 1e3),
                      "write_gbs": gbs, "hbm_peak_gbs": peak, "frac_of_peak": gbs / peak,
                      "spot_check_abs_err": abs(float(out[i, j]) - ref)}))


if __name__ == "__main__":
    main()
