"""world_size-2 (and 4) gloo test of the strip hand-over host logic on CPU: after exchange_halo every rank's halo
rows must equal the neighbours' boundary rows of the full film."""
import os
import socket
import sys

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, height, width, halo, q):
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "vulkan-restir-pt_b200")]
    from restirpt.multigpu import partition, storage_rows, exchange_halo
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    film = torch.arange(height * width * 3, dtype=torch.float32).reshape(height, width, 3)   # "reservoirs" of the full film
    r0, r1 = partition(height, world)[rank]
    s0, s1 = storage_rows(r0, r1, height, halo)
    strip = torch.full((s1 - s0, width, 3), -1.0)
    strip[r0 - s0: r1 - s0] = film[r0:r1]                     # every rank only knows its own rows
    exchange_halo(strip, r0, r1, height, halo, rank, world)
    ok = bool(torch.equal(strip, film[s0:s1]))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,height", [(2, 64), (4, 120)])
def test_halo_exchange_gloo(world, height):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, height, 16, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(results[r] for r in range(world)), results
