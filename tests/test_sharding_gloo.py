"""N>1 host logic on CPU: world_size-2 gloo ranks partition jacobian!/hessian! seed rows / columns exactly as
bench.py does, each rank produces its block (here with the oracle standing in for the device engine — this test checks
the PARTITION and the gather, not the kernels) and an all_gather reassembles the full matrix."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def shard(n, world, rank):
    import __graft_entry__ as graft
    return graft.load_package().shard_range(n, world, rank)


def _worker(rank, world, port, q):
    import __graft_entry__ as graft
    rd = graft.load_package(); orc = graft.load_oracle()
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 12
    A, B, x0 = rd.workloads.inputs_jacobian(n, seed=1)
    tape = rd.JacobianTape(rd.workloads.make_f_jacobian(A, B), x0)
    _, _, x = rd.workloads.inputs_jacobian(n, seed=2)
    rb, re = shard(n, world, rank)
    _, Jblk = orc.OracleTape(tape.program_bytes()).jacobian([x], rows=range(rb, re))     # rows x n, this rank's seeds
    blocks = [torch.zeros(n // world, n, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(blocks, torch.from_numpy(np.ascontiguousarray(Jblk)))
    J = torch.cat(blocks, 0).numpy()
    if rank == 0:
        _, Jfull = orc.OracleTape(tape.program_bytes()).jacobian([x])
        q.put(float(np.max(np.abs(J - Jfull))))
    dist.barrier()
    dist.destroy_process_group()


def test_row_partition_and_gather_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err == 0.0


@pytest.mark.parametrize("n,world", [(8192, 8), (16384, 8), (10, 4), (7, 2)])
def test_shards_cover_exactly(n, world):
    spans = [shard(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a < b
