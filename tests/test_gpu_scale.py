"""GPU (-m gpu): BASELINE.json's full sizes through size-independent properties (symmetry, linearity,
partition-of-work, a row-block sample against the oracle) -- the oracle cannot run n = 434k..2M in full."""
import math

import pytest
import torch

from oracle import cglb_oracle as o

pytestmark = pytest.mark.gpu
f64 = torch.float64


@pytest.fixture(scope="module")
def eng():
    from cglb_b200.engine import get_engine
    return get_engine()


@pytest.mark.parametrize("kind,n,d", [("rbf", 40000, 8), ("matern32", 434000, 3), ("matern32", 515000, 90), ("matern32", 2000000, 11)])
def test_full_size_matvec_properties(eng, kind, n, d):
    dev = eng.device
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=f64, device=dev)
    v = torch.randn(n, generator=g, dtype=f64, device=dev)
    u = torch.randn(n, generator=g, dtype=f64, device=dev)
    ls = torch.full((d,), 0.5 * math.sqrt(d), dtype=f64, device=dev)
    xp = eng.pack(kind, x, ls, x.mean(0))
    kv = eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01)
    # (1) a sample of rows against the CPU oracle (direct-difference form)
    rows = torch.tensor([0, 1, 127, 128, 1023, 1024, n // 2, n - 129, n - 2, n - 1])
    xc, vc, lc = x.cpu(), v.cpu(), ls.cpu()
    ref = o.kernel_dense(kind, xc[rows], xc, lc, torch.tensor(1.0, dtype=f64)) @ vc + 0.01 * vc[rows]
    got = kv[rows.to(dev)].cpu()
    assert float((got - ref).abs().max() / ref.abs().max()) <= 1e-10
    if n <= 500000:
        # (2) symmetry u^T K v = v^T K u and (3) linearity K(v + 2u) = Kv + 2Ku
        ku = eng.kmv_sym(kind, xp, n, d, u, 1.0, 0.01)
        a, b = float(u @ kv), float(v @ ku)
        assert abs(a - b) <= 1e-10 * max(abs(a), abs(b))
        kvu = eng.kmv_sym(kind, xp, n, d, v + 2 * u, 1.0, 0.01)
        assert float((kvu - (kv + 2 * ku)).norm() / kvu.norm()) <= 1e-12
        # (4) the 8-way work partition of the row-sharded run sums to the full product
        parts = sum(eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, part=p, nparts=8) for p in range(8))
        assert float((parts - kv).norm() / kv.norm()) <= 1e-12
