"""The fused feature-row exchange (include/bliss_b200.h: bliss_b200_gather_*): rows stored into every
rank's buffer by finalize_kernel + the one-warp epoch barrier must give every rank exactly the rows a
plain analysis of the whole corpus gives.

* one process, one GPU: two "ranks" (two gather objects connected by raw pointer), barriers on two streams;
* two processes, two GPUs (skipped on a 1-GPU box): real CUDA IPC + NVLink peer stores, compared with
  the NCCL all-gather (tests/gather_worker.py under torch.distributed.run).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import bliss_rs_b200 as B
from bliss_rs_b200 import synth
from bliss_rs_b200.multigpu import _DevArray

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _init():
    B.native.init(0)
    yield


def _corpus(n, dev):
    lengths = [22050 * (4 + (i % 3)) for i in range(n)]
    return synth.gen_corpus_flat(99, list(range(n)), lengths, device=dev), lengths


@pytest.mark.parametrize("version", [2, 1])
def test_two_ranks_in_one_process(version):
    nat = B.native
    dev = torch.device("cuda", 0)
    n, world = 7, 2
    dim = nat.feature_count(version)
    (pcm, offs, lens), _ = _corpus(n, dev)
    want = torch.zeros((n, dim), device=dev)
    st = nat.analyze_batch_device(pcm.data_ptr(), offs, lens, version, want.data_ptr())
    assert (st == 0).all()
    g = [nat.Gather(world, r, n) for r in range(world)]
    try:
        for r in range(world):
            g[r].connect([x.handle for x in g])
            g[r].set_timeout_ms(5000)
        for epoch in range(3):  # both parities of the double buffer, and its reuse
            main = torch.cuda.current_stream()
            for r in range(world):
                ids = list(range(r, n, world))  # round-robin shard
                g[r].scatter(pcm.data_ptr(), [offs[i] for i in ids], [lens[i] for i in ids], version, r, world,
                             None, main.cuda_stream)
            streams = [torch.cuda.Stream(), torch.cuda.Stream()]
            ptrs = []
            for r in range(world):  # the two barriers need each other: they must be on different streams
                streams[r].wait_stream(main)
                ptrs.append(g[r].commit(streams[r].cuda_stream))
            torch.cuda.synchronize()
            for r in range(world):
                g[r].check()
                got = torch.as_tensor(_DevArray(ptrs[r], (n, dim)), device=dev)
                assert torch.equal(got, want), "rank %d epoch %d" % (r, epoch)
    finally:
        for x in g:
            x.destroy()


def test_barrier_times_out_instead_of_hanging():
    nat = B.native
    dev = torch.device("cuda", 0)
    (pcm, offs, lens), _ = _corpus(2, dev)
    g = [nat.Gather(2, r, 2) for r in range(2)]
    try:
        for r in range(2):
            g[r].connect([x.handle for x in g])
        g[0].set_timeout_ms(200)
        g[0].scatter(pcm.data_ptr(), offs[:1], lens[:1], 2, 0, 2)
        g[0].commit()  # rank 1 never commits
        with pytest.raises(nat.NativeError) as e:
            g[0].check()
        assert "timed out waiting for rank 1" in str(e.value)
    finally:
        for x in g:
            x.destroy()


def test_scatter_argument_checks():
    nat = B.native
    dev = torch.device("cuda", 0)
    (pcm, offs, lens), _ = _corpus(3, dev)
    g = nat.Gather(1, 0, 3)
    try:
        with pytest.raises(nat.NativeError):
            g.scatter(pcm.data_ptr(), offs, lens, 2, 0, 1)  # not connected yet
        g.connect([g.handle])
        with pytest.raises(nat.NativeError):
            g.scatter(pcm.data_ptr(), offs, lens, 2, 1, 1)  # rows 1..3 of a 3-row buffer
        with pytest.raises(nat.NativeError):
            g.scatter(pcm.data_ptr(), offs, lens, 2, 0, 0)
        g.scatter(pcm.data_ptr(), offs[:1], lens[:1], 2, 0, 1)
        with pytest.raises(nat.NativeError):
            g.scatter(pcm.data_ptr(), offs[1:], lens[1:], 1, 1, 1)  # versions cannot mix inside an epoch
        g.scatter(pcm.data_ptr(), offs[1:], lens[1:], 2, 1, 1)
        ptr = g.commit()
        g.check()
        want = torch.zeros((3, 23), device=dev)
        nat.analyze_batch_device(pcm.data_ptr(), offs, lens, 2, want.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(torch.as_tensor(_DevArray(ptr, (3, 23)), device=dev), want)
    finally:
        g.destroy()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (CUDA IPC + NVLink peer stores)")
def test_two_processes_match_nccl_all_gather():
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617",
                        os.path.join(ROOT, "tests", "gather_worker.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert "GATHER_WORKER_OK" in r.stdout
