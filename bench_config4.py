#!/usr/bin/env python3
"""BASELINE.json configs[3], literally: 100 000 synthetic 3-min tracks sharded round-robin over the GPUs of one
box, every rank's 23-float rows delivered to every rank, then each rank's row block of the 100 000 x 100 000
distance matrix (src/playlist.rs:140-142 with the v2 weights, src/lib.rs:209-234).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 bench_config4.py

A rank's shard (12 500 tracks = 198 GB of f32 PCM) does not fit its HBM, so the shard is analysed in waves of
--wave tracks whose PCM is generated on the device between the waves (not timed: it stands in for the decoder).
Timed on the device, max over ranks: the analysis of every wave (rows go straight into every rank's row buffer
through peer memory, bliss_b200_analyze_batch_device_scatter; --gather nccl = one all_gather_into_tensor at the
end), the epoch barrier and the distance row block.  One JSON line on rank 0.

Run on 8 B200s in round 2: profiles/config4_r02_8gpu_100k.json (also reachable as `bench.py --config 4`).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
TRACK = 3 * 60 * 22050
BASE_SEED = 20260926


def main(emit=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-songs", type=int, default=100000)
    ap.add_argument("--wave", type=int, default=1024)
    ap.add_argument("--distinct", type=int, default=128, help="distinct generated tracks per wave (the rest are gain-scaled copies)")
    ap.add_argument("--gather", default="auto", choices=["auto", "p2p", "nccl"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bliss_rs_b200 as B
    from bliss_rs_b200 import multigpu as M, synth
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nat = B.native
    nat.init(local_rank)
    n_total, dim = args.total_songs, 23
    mine = M.shard_round_robin(n_total, world, rank)         # global ids rank, rank + world, ...
    n_local = len(mine)
    max_local = -(-n_total // world)
    stream = torch.cuda.current_stream().cuda_stream
    feats = torch.zeros((max_local, dim), dtype=torch.float32, device=dev)
    gather = None
    if world > 1 and args.gather in ("auto", "p2p"):
        ok = torch.ones(1, device=dev)
        try:
            gather = M.PeerGather(n_total, dev)
        except Exception as e:
            if args.gather == "p2p":
                raise
            ok.zero_()
            print("[config4] peer gather unavailable on rank %d: %s" % (rank, e), file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            gather = None

    W = args.wave
    pcm = torch.empty(W * TRACK, dtype=torch.float32, device=dev)
    offs = [i * TRACK for i in range(W)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    analysis_ms, waves = 0.0, 0
    for lo in range(0, n_local, W):
        cnt = min(W, n_local - lo)
        # the wave's PCM (untimed): `distinct` generated tracks, the others gain-scaled copies of them
        d = min(args.distinct, cnt)
        for i in range(d):
            pcm[i * TRACK:(i + 1) * TRACK].copy_(synth.gen_track(BASE_SEED, mine[lo + i], TRACK, dev))
        for i in range(d, cnt):
            g = 0.5 + 0.5 * ((mine[lo + i] * 7919) % 97) / 97.0
            pcm[i * TRACK:(i + 1) * TRACK].copy_(pcm[(i % d) * TRACK:((i % d) + 1) * TRACK] * g)
        torch.cuda.synchronize()
        ev[0].record()
        if gather:  # global row of local song lo + i is rank + (lo + i) * world
            gather.scatter(pcm.data_ptr(), offs[:cnt], [TRACK] * cnt, 2, rank + lo * world, world,
                           feats[lo:lo + cnt].data_ptr(), stream)
        else:
            nat.analyze_batch_device(pcm.data_ptr(), offs[:cnt], [TRACK] * cnt, 2, feats[lo:lo + cnt].data_ptr(), stream)
        ev[1].record()
        torch.cuda.synchronize()
        analysis_ms += ev[0].elapsed_time(ev[1])
        waves += 1
    del pcm
    torch.cuda.empty_cache()

    # exchange + distance row block (timed together; the barrier makes the ranks start it together)
    row_lo, row_hi = M.row_block(n_total, world, rank)
    dmat = torch.empty((row_hi - row_lo, n_total), dtype=torch.float32, device=dev)
    weights = nat.feature_weights(2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev[0].record()
    if gather:
        cols = gather.commit(n_total, dim, stream)
    elif world > 1:
        cols = M.all_gather_features(feats[:n_local], mine, n_total)
    else:
        cols = feats[:n_total]
    rows = cols[row_lo:row_hi]
    nat.distance_matrix_device(rows.data_ptr(), row_hi - row_lo, cols.data_ptr(), n_total, dim, dmat.data_ptr(),
                               nat.METRIC_MAHALANOBIS, weights, stream)
    ev[1].record()
    torch.cuda.synchronize()
    tail_ms = ev[0].elapsed_time(ev[1])
    if gather:
        gather.check()
    # spot check: d(i, i) = 0 on this rank's diagonal entries, rows are the ones this rank computed itself
    diag = dmat[torch.arange(0, row_hi - row_lo, 997, device=dev), torch.arange(row_lo, row_hi, 997, device=dev)]
    own_ok = bool(torch.equal(cols[torch.tensor(mine[:64], device=dev)], feats[:64])) if n_local >= 64 else True
    t = torch.tensor([analysis_ms, tail_ms], dtype=torch.float64, device=dev)
    okt = torch.tensor([float(bool((diag == 0).all()) and own_ok)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        a_ms, d_ms = float(t[0]), float(t[1])
        (emit or (lambda d: print(json.dumps(d), flush=True)))({
            "bench": "configs[3]: %d tracks round-robin over %d GPUs, rows to every rank, all-pairs distance" % (n_total, world),
            "n_gpus": world, "songs": n_total, "waves_per_rank": waves, "wave_songs": W,
            "analysis_ms_max_over_ranks": a_ms, "exchange_plus_distance_ms_max_over_ranks": d_ms,
            "songs_per_s_analysis": n_total / (a_ms / 1e3), "songs_per_s_whole_job": n_total / ((a_ms + d_ms) / 1e3),
            "distance_pairs_per_s": (row_hi - row_lo) * n_total * world / (d_ms / 1e3),
            "exchange": "p2p-fused" if gather else ("nccl" if world > 1 else "none"),
            "checks_ok": bool(okt.item()),
            "note": "PCM generated on the device between waves (untimed); %d distinct tracks per wave" % args.distinct})
    if world > 1:
        dist.barrier()
        if gather:
            gather.destroy()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
