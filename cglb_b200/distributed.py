"""Row sharding of the hot path over the GPUs of one box.

One process per GPU (`torch.distributed`, NCCL over NVLink 5 / NVSwitch).  What is sharded:
  * the work items of the matrix-free sweeps (K v, backward sweep): rank g processes the tile pairs
    {t : t % world == g} of the symmetric enumeration (square blocks in csrc/kmv_impl.cuh, row-block x
    column-chunk strips in csrc/dsweep_impl.cuh; `symmetric_items` / `strip_items` below) and the partial n-vectors are
    all-reduced -- with the symmetric sweep a tile contributes to two row blocks, so the exchange is an
    all-reduce of y rather than an all-gather of p;
  * the columns of the dense M x n matrix A = L^-1 K_uf / sigma (rows of K_nm): rank g owns the
    contiguous column block [lo_g, hi_g).
Replicated: X, y, Z, every n-vector of the solver (16 MB at n = 2M), all M x M factors.  Because every
all-reduced quantity is bit-identical on all ranks, all ranks take the same CG branches.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Optional, Tuple

import torch

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


@dataclass
class Shard:
    rank: int = 0
    world: int = 1
    group: Any = None

    @staticmethod
    def from_env() -> "Shard":
        """The default process group if torch.distributed is initialised, else a single shard."""
        if dist is not None and dist.is_available() and dist.is_initialized():
            return Shard(dist.get_rank(), dist.get_world_size(), None)
        return Shard()

    # ---- partition arithmetic (pure python; covered by the CPU gloo tests) --------------------------
    def column_block(self, n: int) -> Tuple[int, int]:
        """Contiguous, even-aligned column block [lo, hi) of an n-column matrix owned by this rank."""
        per = -(-n // self.world)
        per += per & 1
        lo = min(self.rank * per, n)
        hi = min(lo + per, n)
        return lo, hi

    def owns_item(self, t: int) -> bool:
        """Same rule as the kernels: global work item t belongs to part t % world."""
        return t % self.world == self.rank

    # ---- collectives ---------------------------------------------------------------------------------
    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def broadcast(self, t: torch.Tensor, src: int = 0) -> torch.Tensor:
        if self.world > 1:
            dist.broadcast(t, src=src, group=self.group)
        return t

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)


def symmetric_items(n: int, block: int):
    """Enumeration of the symmetric sweep's work items: t = C(C+1)/2 + I for row block I <= column
    block C (must match item_to_blocks_sym in csrc/kmv_impl.cuh)."""
    nb = -(-n // block)
    t = 0
    for c in range(nb):
        for i in range(c + 1):
            yield t, i, c
            t += 1


def decode_strip_item(t: int, n_chunks: int, rows_per_chunk: int, superrow_chunks: int):
    """Item t of the DMMA sweeps -> (row block I, column chunk c): the host mirror of DCursor::decode in
    csrc/dsweep_impl.cuh.  Row blocks are grouped into super-rows of R = rows_per_chunk * Q blocks (Q = superrow_chunks);
    super-row s pairs its blocks with the chunks c >= s Q, chunk-major: rows_per_chunk (k + 1) items for the local chunk
    k = c - s Q < Q (triangular part), R items for k >= Q."""
    rpc, q = rows_per_chunk, superrow_chunks
    r = rpc * q
    s = 0
    while True:
        nc = n_chunks - s * q
        kk = min(nc, q)
        cnt = rpc * kk * (kk + 1) // 2 + (r * (nc - q) if nc > q else 0)
        if t < cnt or nc <= q:
            break
        t -= cnt
        s += 1
    tri = rpc * q * (q + 1) // 2
    if t < tri:
        k = int(((1.0 + 8.0 * t / rpc) ** 0.5 - 1.0) * 0.5)
        while rpc * k * (k + 1) // 2 > t:
            k -= 1
        while rpc * (k + 1) * (k + 2) // 2 <= t:
            k += 1
        iloc = t - rpc * k * (k + 1) // 2
    else:
        k = q + (t - tri) // r
        iloc = (t - tri) % r
    return s * r + iloc, s * q + k


def strip_items(n: int, rows: int = 256, rows_per_chunk: int = 4, tile: int = 64, superrow_chunks: int = 1 << 30):
    """Enumeration of the DMMA sweeps' work items (DCursor in csrc/dsweep_impl.cuh): item t pairs the row block I (`rows`
    rows) with the column chunk C (`rows_per_chunk * rows` columns), for every I < rows_per_chunk (C + 1) that exists, in the
    L2-blocked order of `decode_strip_item`; only the columns at or after the row block are visited, in tiles of `tile`
    columns.  Yields (t, r0, r1, [(j0, j1, offdiag), ...]): tiles overlapping the row block (offdiag False) are evaluated as
    ordered pairs and feed the row sums only, tiles beyond it feed the row AND the column sums."""
    chunk = rows_per_chunk * rows
    nb_rows = -(-n // rows)
    n_chunks = -(-n // chunk)
    nitems = rows_per_chunk * n_chunks * (n_chunks + 1) // 2
    for t in range(nitems):
        i, c = decode_strip_item(t, n_chunks, rows_per_chunk, superrow_chunks)
        if i >= nb_rows or c >= n_chunks:
            continue
        r0 = i * rows
        cbeg, cend = max(c * chunk, r0), min((c + 1) * chunk, n)
        if cbeg >= cend:
            continue
        tiles = [(j0, min(j0 + tile, cend), j0 >= r0 + rows) for j0 in range(cbeg, cend, tile)]
        yield t, r0, min(r0 + rows, n), tiles
