"""Host-side logic of the multi-GPU path.

Songs are independent units, so the analysis shards with NO data-path collective: song i goes
to rank `i % world` (BASELINE.json config 4, "sharded round-robin") or, for mixed durations
(config 5), to the currently least-loaded rank in longest-first order.  The one exchange step is
the all-gather of the [n_local x dim] f32 feature rows before the all-pairs distance; afterwards
rank r owns the row block `row_block(n, world, r)` of the n x n matrix.

Works on any torch.distributed backend: NCCL on the GPU box, gloo in the CPU tests.

`PeerGather` is the fused form of that exchange on one NVSwitch box: the last kernel of the analysis
stores each finished row into every rank's row buffer (peer-mapped memory), so the only thing left of
the collective is a one-warp epoch barrier (include/bliss_b200.h, "fused feature-row exchange").
"""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_round_robin(n_songs: int, world: int, rank: int) -> List[int]:
    """global song indices owned by `rank`: i with i % world == rank (config 4)"""
    return list(range(rank, n_songs, world))


def shard_longest_first(lengths: Sequence[int], world: int) -> List[List[int]]:
    """LPT greedy: songs in decreasing length to the least-loaded rank (config 5).
    Returns one index list per rank; each list is sorted by global index."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    return [sorted(v) for v in out]


def row_block(n: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) rows of the n x n distance matrix computed by `rank` (contiguous, balanced)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_features(local: torch.Tensor, owned: Sequence[int], n_total: int, group=None) -> torch.Tensor:
    """All ranks end up with the full [n_total, dim] matrix in GLOBAL song order.

    local: [len(owned), dim] rows of this rank (device or CPU tensor); owned: their global indices.
    Shards may have different sizes: rows are padded to the largest shard for the collective."""
    world = dist.get_world_size(group)
    dim = local.shape[1]
    counts = [None] * world
    dist.all_gather_object(counts, [int(i) for i in owned], group=group)
    m = max(len(c) for c in counts)
    padded = torch.zeros((m, dim), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    gathered = torch.empty((world * m, dim), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    full = torch.empty((n_total, dim), dtype=local.dtype, device=local.device)
    for r, idx in enumerate(counts):
        if idx:
            full[torch.tensor(idx, device=local.device)] = gathered[r * m: r * m + len(idx)]
    return full


def round_robin_to_global(gathered: torch.Tensor, world: int) -> torch.Tensor:
    """Equal-size round-robin shards: [world*S, dim] in rank-major order -> global order
    (row i*world + r = row i of rank r).  Pure view/permutation, no index tensors."""
    s = gathered.shape[0] // world
    return gathered.view(world, s, -1).transpose(0, 1).reshape(world * s, -1)


class _DevArray:
    """zero-copy torch view of a raw device pointer (via __cuda_array_interface__)"""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerGather:
    """Fused all-gather of feature rows over peer memory (one process per GPU, one box).

    Construction is collective: every rank creates its buffer, the 128-byte handles travel through
    `dist.all_gather_object`, every rank maps every peer.  Per step: `scatter(...)` (the analysis;
    global row of local song i = row_offset + i * row_stride), then `commit()` -> [n_rows, dim] view of
    this rank's complete row buffer, valid until the commit after next."""

    def __init__(self, max_rows: int, device, group=None):
        from . import _native as nat
        self._nat = nat
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.max_rows = int(max_rows)
        self.g = nat.Gather(self.world, self.rank, self.max_rows)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.g.handle, group=group)
        self.g.connect(handles)

    def scatter(self, d_pcm_ptr, offsets, n_samples, version, row_offset, row_stride, d_out_ptr=None,
                stream_ptr=None):
        return self.g.scatter(d_pcm_ptr, offsets, n_samples, version, row_offset, row_stride, d_out_ptr,
                              stream_ptr)

    def commit(self, n_rows: int, dim: int, stream_ptr=None) -> torch.Tensor:
        ptr = self.g.commit(stream_ptr)
        return torch.as_tensor(_DevArray(ptr, (int(n_rows), int(dim))), device=self.device)

    def check(self):
        self.g.check()

    def destroy(self):
        self.g.destroy()
