"""Launched by test_gpu_gather.py / scripts/gpu_multi2.sh under torch.distributed.run: every rank analyses
its round-robin shard, rows travel (a) by the fused peer stores + epoch barrier and (b) by the NCCL
all-gather; both must give every rank the rows of a whole-corpus analysis, bit for bit."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bliss_rs_b200 as B  # noqa: E402
from bliss_rs_b200 import multigpu as M, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nat = B.native
    nat.init(local)
    per_rank = 5
    n = per_rank * world
    dim = 23
    lengths = [22050 * (3 + (i % 4)) for i in range(n)]
    ids = M.shard_round_robin(n, world, rank)
    pcm, offs, lens = synth.gen_corpus_flat(7, ids, [lengths[i] for i in ids], device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    pg = M.PeerGather(n, dev)
    feats = torch.zeros((per_rank, dim), device=dev)
    for epoch in range(4):
        pg.scatter(pcm.data_ptr(), offs, lens, 2, rank, world, feats.data_ptr(), stream)
        cols = pg.commit(n, dim, stream)
        fused = cols.clone()
        gathered = torch.zeros((n, dim), device=dev)
        dist.all_gather_into_tensor(gathered, feats)
        nccl = M.round_robin_to_global(gathered, world)
        torch.cuda.synchronize()
        pg.check()
        assert torch.equal(fused, nccl), "rank %d epoch %d: fused rows differ from the NCCL all-gather" % (rank, epoch)
        assert torch.equal(fused[rank::world], feats)
    # rank 0 also analyses the whole corpus on its own: global order, same bits
    if rank == 0:
        apcm, aoffs, alens = synth.gen_corpus_flat(7, list(range(n)), lengths, device=dev)
        want = torch.zeros((n, dim), device=dev)
        nat.analyze_batch_device(apcm.data_ptr(), aoffs, alens, 2, want.data_ptr(), stream)
        torch.cuda.synchronize()
        assert torch.equal(fused, want)
    dist.barrier()
    pg.destroy()
    if rank == 0:
        print("GATHER_WORKER_OK world=%d rows=%d" % (world, n), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
