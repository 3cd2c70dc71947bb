"""N>1 host logic on CPU: two gloo ranks, one parameter broadcast, utterance
shards that tile the control file exactly, hypotheses gathered in order."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from cmusphinx_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    params = None
    if rank == 0:
        rng = np.random.default_rng(0)
        params = dict(mean=rng.standard_normal((7, 4, 5)).astype(np.float32),
                      var=rng.random((7, 4, 5)).astype(np.float32),
                      det=rng.standard_normal((7, 4)).astype(np.float32),
                      mixw=rng.integers(0, 255, (7, 1, 4)).astype(np.uint8))
    got, digest = shard.broadcast_params(params, src=0)
    n_utt = 11
    mine = list(shard.shard_strided(n_utt, rank, world))
    hyps = shard.gather_hypotheses([(i, f"utt{i}") for i in mine], dst=0)
    q.put((rank, digest, {k: v.shape for k, v in got.items()}, float(got["mean"].sum()), mine, hyps))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_broadcast_shards_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, d0, s0, m0, u0, h0), (r1, d1, s1, m1, u1, h1) = res
    assert d0 == d1 and s0 == s1 and m0 == m1           # identical parameter blobs after the one broadcast
    assert sorted(u0 + u1) == list(range(11)) and not set(u0) & set(u1)
    assert h0 == [f"utt{i}" for i in range(11)] and h1 is None


@pytest.mark.parametrize("n,world", [(0, 4), (1, 4), (10, 3), (100000, 8), (7, 8)])
def test_shards_tile_the_control_file(n, world):
    for fn in (shard.shard_strided, shard.shard_block):
        parts = [list(fn(n, r, world)) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip():
    rng = np.random.default_rng(1)
    p = dict(mean=rng.standard_normal((3, 2, 5)).astype(np.float32), var=rng.random((3, 2, 5)).astype(np.float32),
             det=rng.standard_normal((3, 2)).astype(np.float32), mixw=rng.integers(0, 255, (3, 1, 2)).astype(np.uint8))
    blob, man = shard.pack_params(p)
    q = shard.unpack_params(blob, man)
    for k in p:
        np.testing.assert_array_equal(p[k], q[k])
