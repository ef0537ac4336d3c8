"""CPU tests of the multi-GPU host logic: the x-strip partition and the neighbour
halo exchange protocol of sem2dpack_b200.strips, run over the gloo backend with world_size 2 and 3."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sem2dpack_b200 import strips


def test_partition_covers_the_box():
    for nx, world in [(8192, 8), (30, 3), (31, 4), (7, 7), (10, 1)]:
        p = strips.partition(nx, world)
        assert p[0][0] == 0 and p[-1][1] == nx
        assert all(p[r][1] == p[r + 1][0] for r in range(world - 1))
        w = [hi - lo for lo, hi in p]
        assert max(w) - min(w) <= 1 and min(w) >= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nz, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank holds partial sums of its two interface columns: value = f(rank, side, row)
        def col(r, side):
            return torch.arange(nz, dtype=torch.float64) * 0.5 + 100.0 * r + 7.0 * side

        send_l = col(rank, 0) if rank > 0 else None
        send_r = col(rank, 1) if rank < world - 1 else None
        recv_l = torch.zeros(nz, dtype=torch.float64) if rank > 0 else None
        recv_r = torch.zeros(nz, dtype=torch.float64) if rank < world - 1 else None
        for _ in range(3):  # repeated steps reuse the same buffers
            for req in strips.exchange_halos(send_l, send_r, recv_l, recv_r, rank, world):
                req.wait()
        ok = True
        if rank > 0:
            ok &= bool(torch.equal(recv_l, col(rank - 1, 1)))
            # own + neighbour's is the same number on both sides of the interface
            total_here = send_l + recv_l
            total_there = col(rank - 1, 1) + col(rank, 0)
            ok &= bool(torch.equal(total_here, total_there))
        if rank < world - 1:
            ok &= bool(torch.equal(recv_r, col(rank + 1, 0)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_neighbour_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 257, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=10) for _ in range(world))
    assert got == [(r, True) for r in range(world)]
