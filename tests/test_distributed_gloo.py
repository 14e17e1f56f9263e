"""CPU, world_size 2, gloo: the N>1 host path.  Each rank evaluates only its share of the symmetric work
items / its column block of A (numpy emulation of what the kernels do for that share, using the product's
own partition helpers), exchanges through `Shard` exactly as bound.py / conjugate_gradient.py do, and the
result must equal the single-rank oracle."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cglb_b200.distributed import Shard, symmetric_items
from oracle import cglb_oracle as o


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = Shard.from_env()
        assert (shard.rank, shard.world) == (rank, world)
        n, d, m, block = 300, 2, 10, 64
        x, y, z = o.synthetic_problem(n, d, m, seed=3)
        ls = torch.tensor([[0.8, 1.3]], dtype=torch.float64)
        var = torch.tensor(1.2, dtype=torch.float64)
        v = torch.randn(n, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(9))
        K = o.kernel_dense("matern32", x, x, ls, var)
        # symmetric sweep share of this rank: tile (I, C), I <= C contributes to y_I and (off-diagonal) y_C
        ypart = torch.zeros(n, 1, dtype=torch.float64)
        for t, i, c in symmetric_items(n, block):
            if not shard.owns_item(t):
                continue
            rs, cs = slice(i * block, min(n, (i + 1) * block)), slice(c * block, min(n, (c + 1) * block))
            tile = K[rs, cs]
            ypart[rs] += tile @ v[cs]
            if i != c:
                ypart[cs] += tile.T @ v[rs]
        if rank == 0:
            ypart += 0.1 * v                                    # the diag term is added by part 0
        shard.all_reduce(ypart)
        ok1 = torch.allclose(ypart, K @ v + 0.1 * v, rtol=1e-12, atol=1e-12)
        # preconditioner: column block of A, q all-reduced, (z, rz) all-reduced in one buffer
        terms = o.common_terms("matern32", o.OracleParams.from_values(0.1, 0.0, z, 1.2, [0.8, 1.3]), x, 1e-6)
        A, LB = terms.A.detach(), terms.LB.detach()
        lo, hi = shard.column_block(n)
        r = torch.randn(n, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(11))
        q = A[:, lo:hi] @ r[lo:hi]
        shard.all_reduce(q)
        w = torch.cholesky_solve(q, LB)
        zbuf = torch.zeros(n + 1, dtype=torch.float64)
        zbuf[lo:hi] = ((r[lo:hi] - A[:, lo:hi].T @ w) / 0.1).reshape(-1)
        zbuf[n] = (zbuf[lo:hi] * r[lo:hi].reshape(-1)).sum()
        shard.all_reduce(zbuf)
        zref, rzref = o.nystrom_preconditioner(A, LB, torch.tensor(0.1, dtype=torch.float64))(r)
        ok2 = torch.allclose(zbuf[:n], zref.reshape(-1), rtol=1e-10, atol=1e-12) and abs(float(zbuf[n]) - float(rzref)) < 1e-9 * abs(float(rzref))
        flag = torch.tensor([float(ok1 and ok2)])
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            ret.put(bool(flag.item()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_matvec_and_preconditioner_match_single_rank():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True
