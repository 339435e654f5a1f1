"""N > 1 host logic on CPU: world_size-2 gloo processes shard the images, build their
slice of a score table and exchange it with the path's single all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scanpaths_b200.dist import allgather_tables, padded_shard, shard_range


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    K = 3
    full = torch.arange(K * n_total * 11, dtype=torch.float32).view(K, n_total, 11)
    mine = full[:, lo:hi].clone()
    got = allgather_tables(mine, n_total, image_dim=1)
    ok = torch.equal(got, full)
    q.put((rank, ok, tuple(got.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7, 1])
def test_sharded_tables_allgather_gloo(n_total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok and shape == (3, n_total, 11), (rank, ok, shape)


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 4096, 4099):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in r) <= padded_shard(n, w)
