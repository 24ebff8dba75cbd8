"""world_size-2 gloo tests (CPU) of the host logic of the N>1 path: sharding, the feature
all-gather into global order, and the distance-matrix row blocks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bliss_rs_b200 import multigpu as M


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _feat(i, dim=23):
    return torch.arange(dim, dtype=torch.float32) * 0.01 + float(i)


def _worker(rank, world, port, n_songs, lengths):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # config 4: round robin, equal shards -> pure permutation path
        own = M.shard_round_robin(n_songs, world, rank)
        local = torch.stack([_feat(i) for i in own])
        full = M.all_gather_features(local, own, n_songs)
        want = torch.stack([_feat(i) for i in range(n_songs)])
        assert torch.equal(full, want)
        if n_songs % world == 0:
            g = torch.empty((n_songs, 23))
            dist.all_gather_into_tensor(g, local)
            assert torch.equal(M.round_robin_to_global(g, world), want)
        # config 5: mixed durations, LPT shards of different sizes
        shards = M.shard_longest_first(lengths, world)
        own = shards[rank]
        local = torch.stack([_feat(i) for i in own]) if own else torch.zeros((0, 23))
        full = M.all_gather_features(local, own, len(lengths))
        assert torch.equal(full, torch.stack([_feat(i) for i in range(len(lengths))]))
        # row blocks tile the matrix exactly once
        lo, hi = M.row_block(len(lengths), world, rank)
        block = torch.cdist(full[lo:hi], full)
        blocks = [None] * world
        dist.all_gather_object(blocks, (lo, hi, block.numpy()))
        if rank == 0:
            rows = sorted(blocks)
            assert rows[0][0] == 0 and rows[-1][1] == len(lengths)
            for a, b in zip(rows, rows[1:]):
                assert a[1] == b[0]
            whole = np.concatenate([b[2] for b in rows])
            assert np.allclose(whole, torch.cdist(full, full).numpy())
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharding_and_gather():
    lengths = [661500, 13230000, 3969000, 700000, 9000000, 661500, 5000000]
    mp.spawn(_worker, args=(2, _free_port(), 10, lengths), nprocs=2, join=True)


def test_shard_properties():
    for world in (1, 2, 4, 8):
        all_idx = sorted(i for r in range(world) for i in M.shard_round_robin(37, world, r))
        assert all_idx == list(range(37))
        blocks = [M.row_block(37, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == 37
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    rng = np.random.default_rng(0)
    lengths = (rng.zipf(1.5, 200).clip(1, 20) * 661500).tolist()
    shards = M.shard_longest_first(lengths, 8)
    assert sorted(i for s in shards for i in s) == list(range(200))
    loads = [sum(lengths[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(lengths)  # LPT bound
