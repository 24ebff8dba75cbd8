"""Where do the fused and the NCCL exchange differ?  torchrun --nproc-per-node 2 scripts/gather_debug.py [songs_per_gpu]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bliss_rs_b200 as B  # noqa: E402
from bliss_rs_b200 import multigpu as M, synth  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nat = B.native
nat.init(local)
n, dim = S * world, 23
ids = M.shard_round_robin(n, world, rank)
pcm, offs, lens = synth.gen_corpus_flat(20240917, ids, [3969000] * S, device=dev)
stream = torch.cuda.current_stream().cuda_stream
feats = torch.zeros((S, dim), device=dev)
runs = []
for i in range(3):
    nat.analyze_batch_device(pcm.data_ptr(), offs, lens, 2, feats.data_ptr(), stream)
    torch.cuda.synchronize()
    runs.append(feats.clone())
for i in (1, 2):
    d = (runs[i] != runs[0])
    print("rank %d: plain run %d vs 0: %d differing values, rows %s cols %s" % (
        rank, i, int(d.sum()), d.any(1).nonzero().flatten().tolist()[:8], d.any(0).nonzero().flatten().tolist()), flush=True)
if world > 1:
    pg = M.PeerGather(n, dev)
    for ep in range(3):
        pg.scatter(pcm.data_ptr(), offs, lens, 2, rank, world, feats.data_ptr(), stream)
        cols = pg.commit(n, dim, stream).clone()
        torch.cuda.synchronize()
        g = torch.zeros((n, dim), device=dev)
        dist.all_gather_into_tensor(g, feats)
        nc = M.round_robin_to_global(g, world)
        torch.cuda.synchronize()
        d = cols != nc
        own = (cols[rank::world] != feats)
        print("rank %d epoch %d: fused vs nccl %d differing values (rows %s, cols %s); own rows vs local out: %d; local out vs run0: %d" % (
            rank, ep, int(d.sum()), d.any(1).nonzero().flatten().tolist()[:8], d.any(0).nonzero().flatten().tolist(),
            int(own.sum()), int((feats != runs[0]).sum())), flush=True)
        if d.any():
            r = int(d.any(1).nonzero().flatten()[0])
            print("rank %d row %d fused %s\n nccl %s" % (rank, r, cols[r].tolist(), nc[r].tolist()), flush=True)
    pg.check()
    dist.barrier()
    pg.destroy()
    dist.destroy_process_group()
