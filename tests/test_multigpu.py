"""Two ranks, two GPUs (-m gpu, skipped on a single-GPU box): jacobian! row shards and hessian! column shards computed per rank
and assembled by the product's own gather (rdb_gather_result: NCCL inside librdb200.so) into the reference's column-major
result (src/api/utils.jl:44-58), with an UNEVEN partition; every rank must hold the matrix a single GPU computes."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)          # bootstrap only: the data moves over rdb_comm's NCCL
    torch.cuda.set_device(rank)
    rd = graft.load_package()
    ctx = rd.engine.Context(rank)
    comm = rd.make_comm(ctx)
    out = {}
    # ---- jacobian!: n = 515 (uneven: 257 + 258 rows), the int8-slice GEMMs are in play (>= 256 seeds per block)
    n = 515
    A, B, x0 = rd.workloads.inputs_jacobian(n, seed=1)
    tape = rd.JacobianTape(rd.workloads.make_f_jacobian(A, B), x0)
    ct = rd.compile(tape, ctx=ctx)
    _, _, x = rd.workloads.inputs_jacobian(n, seed=2)
    rb, re = rd.shard_range(n, world, rank)
    blk = torch.empty((re - rb) * n, dtype=torch.float64, device="cuda")
    rd.jacobian_(blk, ct, torch.from_numpy(x).cuda(), rows=(rb, re))
    J = torch.full((n * n,), float("nan"), dtype=torch.float64, device="cuda")
    rd.gather_jacobian_(J, blk, ct, comm)
    s = A @ x + x; t = np.tanh(B @ x)
    Jcf = np.diag(t) @ (A + np.eye(n)) + np.diag(s * (1 - t ** 2)) @ B
    Jh = J.cpu().numpy().reshape(n, n, order="F")
    out["jac"] = float(np.max(np.abs(Jh - Jcf)) / np.max(np.abs(Jcf)))
    # gather to rank 1 only
    J1 = torch.full((n * n,), float("nan"), dtype=torch.float64, device="cuda") if rank == 1 else None
    rd.gather_jacobian_(J1, blk, ct, comm, root=1)
    out["jac_root"] = float(np.max(np.abs(J1.cpu().numpy().reshape(n, n, order="F") - Jh))) if rank == 1 else 0.0
    # ---- hessian!: extended Rosenbrock n = 1002 (501 columns per rank), zeros exact
    n = 1002
    ht = rd.compile(rd.HessianTape(rd.workloads.f_rosenbrock, rd.workloads.inputs_rosenbrock(n, seed=1)), ctx=ctx)
    xh = rd.workloads.inputs_rosenbrock(n, seed=2)
    cb, ce = rd.shard_range(n, world, rank)
    hb = torch.empty(n * (ce - cb), dtype=torch.float64, device="cuda")
    rd.hessian_(hb, ht, torch.from_numpy(xh).cuda(), cols=(cb, ce))
    H = torch.full((n * n,), float("nan"), dtype=torch.float64, device="cuda")
    rd.gather_hessian_(H, hb, ht, comm)
    o, e = xh[0::2], xh[1::2]
    Hcf = np.zeros((n, n)); i = np.arange(n // 2)
    Hcf[2 * i, 2 * i] = 1200 * o ** 2 - 400 * e + 2
    Hcf[2 * i, 2 * i + 1] = Hcf[2 * i + 1, 2 * i] = -400 * o
    Hcf[2 * i + 1, 2 * i + 1] = 200
    Hh = H.cpu().numpy().reshape(n, n, order="F")
    out["hess"] = float(np.max(np.abs(Hh - Hcf)) / np.max(np.abs(Hcf)))
    out["hess_zeros"] = bool(np.all(Hh[Hcf == 0] == 0.0))
    q.put((rank, out))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_two_gpu_gather_into_reference_layout():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under `gpurun --gpus 2`)")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in (0, 1):
        assert got[r]["jac"] < 1e-12 and got[r]["hess"] < 1e-12 and got[r]["hess_zeros"], got
    assert got[1]["jac_root"] == 0.0
