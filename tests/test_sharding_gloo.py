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
    # what a rank's rdb_jacobian(row_begin, row_end, ld = rows) leaves behind: a contiguous column-major (rows x n) block
    mine = torch.from_numpy(np.ascontiguousarray(np.asfortranarray(Jblk).reshape(-1, order="F")))
    bounds = rd.shard_bounds(n, world)
    blocks = [torch.zeros((bounds[r + 1] - bounds[r]) * n, dtype=torch.float64) for r in range(world)]
    dist.all_gather(blocks, mine)                                    # gloo stands in for NCCL; the PLACEMENT is the product's
    J = np.full((n, n), np.nan, order="F")
    for r in range(world):                                           # rdb_place_block: the host twin of rdb_gather_result
        rows = bounds[r + 1] - bounds[r]
        rd.engine.place_block(0, blocks[r].numpy().reshape(rows, n, order="F"), bounds[r], bounds[r + 1], J)
    # column blocks (hessian! shards): contiguous in the column-major result
    Hfull = np.arange(float(n * n)).reshape(n, n, order="F")
    H = np.full((n, n), np.nan, order="F")
    for r in range(world):
        rd.engine.place_block(1, Hfull[:, bounds[r]:bounds[r + 1]], bounds[r], bounds[r + 1], H)
    if rank == 0:
        _, Jfull = orc.OracleTape(tape.program_bytes()).jacobian([x])
        q.put(max(float(np.max(np.abs(J - Jfull))), float(np.max(np.abs(H - Hfull)))))
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
