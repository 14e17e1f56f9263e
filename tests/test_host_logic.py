"""CPU: host-side logic that mirrors the reference interface (parameter order and transforms, SciPy
pack/unpack, configs, inducing-point initialisation, shard arithmetic)."""
import numpy as np
import pytest
import torch

import cglb_b200 as cb
from cglb_b200.distributed import Shard, symmetric_items
from cglb_b200.inducing import ConditionalVariance
from helpers import make_model
from oracle import cglb_oracle as o


def test_parameter_order_shapes_and_transforms():
    x, y, z = o.synthetic_problem(40, 3, 6)
    m = make_model("matern32", x.numpy(), y.numpy(), z.numpy(), 0.25, 1.7, [0.5, 1.0, 2.0], 0.3, device="cpu")
    names = [n for n, _ in m.named_parameters()]
    assert names == ["likelihood.noise_covar.raw_noise", "mean_module.constant", "covar_module.inducing_points",
                     "covar_module.base_kernel.raw_outputscale", "covar_module.base_kernel.base_kernel.raw_lengthscale"]
    shapes = [tuple(p.shape) for p in m.parameters()]
    assert shapes == [(1,), (1,), (6, 3), (), (1, 3)]
    assert abs(float(m.likelihood.noise) - 0.25) < 1e-12          # softplus(raw) + 1e-6
    assert abs(float(m.covar_module.base_kernel.outputscale) - 1.7) < 1e-12
    assert np.allclose(m.covar_module.base_kernel.base_kernel.lengthscale.detach().numpy(), [[0.5, 1.0, 2.0]])
    p = o.OracleParams.from_values(0.25, 0.3, z, 1.7, [0.5, 1.0, 2.0])
    for a, b in zip(m.parameters(), p.tensors()):
        assert np.allclose(a.detach().numpy(), b.detach().numpy(), atol=1e-12)
    assert m.v_vec.shape == (40, 1) and not m.v_vec.requires_grad
    pars = cb.interface.model_parameters(m) if hasattr(cb, "interface") else None
    from cglb_b200.interface import model_parameters
    pars = model_parameters(m)
    assert set(pars) == {".likelihood.variance", ".mean_function.c", ".inducing_variable.Z", ".kernel.lengthscales", ".kernel.variance"}


def test_scipy_pack_unpack_assign_roundtrip():
    ts = [torch.zeros(1, dtype=torch.float64), torch.zeros((), dtype=torch.float64), torch.zeros(2, 3, dtype=torch.float64)]
    vec = torch.arange(8, dtype=torch.float64)
    vals = cb.Scipy.unpack(ts, vec)
    assert [tuple(v.shape) for v in vals] == [(1,), (), (2, 3)]
    cb.Scipy.assign(ts, vals)
    assert torch.equal(cb.Scipy.pack(ts), vec)
    with pytest.raises(ValueError):
        cb.Scipy.assign(ts, vals[:2])


def test_configs_and_registries():
    data = (np.zeros((5, 4)), np.zeros((5, 1)))
    assert cb.KERNEL_CONFIGS["mat32"] is cb.Matern32Config and cb.KERNEL_CONFIGS["rbf"] is cb.SquaredExponentialConfig
    kp = cb.Matern32Config().params(data)
    assert kp["variance"] == 1.0 and np.allclose(kp["lengthscales"], np.ones(4))
    cfg = cb.CGLBConfig(kernel=cb.Matern32Config(), inducing_variable=cb.InducingVariableConfig(3))
    p = cfg.params(data)
    assert p["noise_variance"] == 1.0 and p["max_error"] == 1.0 and callable(p["inducing_variable"])
    with pytest.raises(Exception):
        cfg.max_error = 2.0                                    # frozen
    from cglb_b200 import interface
    with pytest.raises(NotImplementedError):
        interface.set_default_float("fp16")
    with pytest.raises(NotImplementedError):
        interface.create_model(object(), data)
    assert set(cb.BACKENDS) >= {"b200", "torch"}
    cb.B200.set_default_jitter("fp32")
    from cglb_b200 import settings
    assert settings.cholesky_jitter.value() == 1e-5
    cb.B200.set_default_jitter("fp64")
    assert settings.cholesky_jitter.value() == 1e-6


def test_conditional_variance_greedy_matches_bruteforce():
    rng = np.random.RandomState(0)
    X = rng.randn(60, 2)
    ls = 0.7

    def kern(x1, x2, full_cov=False):
        if not full_cov:
            return np.ones(x1.shape[0])
        x2 = x1 if x2 is None else x2
        d2 = ((x1[:, None, :] - x2[None, :, :]) ** 2).sum(-1)
        return np.exp(-0.5 * d2 / ls ** 2)

    Z, idx = ConditionalVariance(sample=False)(X, 8, kern)
    assert Z.shape == (8, 2) and len(set(idx.tolist())) == 8 and np.allclose(X[idx], Z)
    # brute force: each new point maximises the conditional variance given the already chosen ones
    chosen = [idx[0]]
    for step in range(1, 8):
        Kzz = kern(X[chosen], X[chosen], True) + 1e-12 * np.eye(len(chosen))
        Kxz = kern(X, X[chosen], True)
        cond = 1.0 - np.einsum("ij,jk,ik->i", Kxz, np.linalg.inv(Kzz), Kxz)
        assert cond[idx[step]] >= cond.max() - 1e-9
        chosen.append(idx[step])


def test_shard_column_blocks_partition_and_items():
    for n in (1, 7, 300, 2001):
        for world in (1, 2, 3, 8):
            blocks = [Shard(r, world).column_block(n) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (l0, h0), (l1, h1) in zip(blocks, blocks[1:]):
                assert h0 == l1 and (l1 % 2 == 0 or l1 == h1)      # non-empty blocks start 16-byte aligned
    items = list(symmetric_items(1000, 256))
    assert len(items) == 10 and all(i <= c for _, i, c in items)
    assert [t for t, _, _ in items] == list(range(10))
    owned = [sum(Shard(r, 3).owns_item(t) for r in range(3)) for t, _, _ in items]
    assert owned == [1] * 10


def test_strip_items_cover_every_pair_once():
    """The DMMA sweeps' decomposition (strip_items mirrors DCursor in dsweep_impl.cuh): emulating what the kernel does
    with each item -- ordered pairs with row sums only on the tiles overlapping the row block, row AND column sums
    beyond it -- reproduces K v exactly, for any split of the items over ranks."""
    import numpy as np
    from cglb_b200.distributed import Shard, strip_items
    rng = np.random.default_rng(0)
    for n, rows, rpc, tile, q in [(700, 64, 4, 16, 1 << 30), (1030, 128, 4, 64, 1), (257, 256, 4, 64, 2), (64, 16, 2, 8, 1),
                                  (999, 32, 4, 8, 3), (999, 32, 4, 8, 2), (2050, 16, 4, 16, 5), (1500, 16, 4, 16, 7)]:
        a = rng.standard_normal((n, n))
        K = a + a.T                                   # any symmetric matrix
        v = rng.standard_normal(n)
        ref = K @ v
        for world in (1, 3):
            y = np.zeros(n)
            seen = set()
            for rank in range(world):
                sh = Shard(rank, world)
                for t, r0, r1, tiles in strip_items(n, rows, rpc, tile, superrow_chunks=q):
                    if not sh.owns_item(t):
                        continue
                    assert t not in seen
                    seen.add(t)
                    for j0, j1, offdiag in tiles:
                        blk = K[r0:r1, j0:j1]
                        y[r0:r1] += blk @ v[j0:j1]
                        if offdiag:
                            y[j0:j1] += blk.T @ v[r0:r1]
                        else:
                            assert j0 >= r0 and j1 <= r0 + rows      # inside the diagonal block: ordered pairs
            assert np.allclose(y, ref, rtol=1e-12, atol=1e-10), (n, rows, world, q)


def test_strip_item_order_is_a_bijection_and_l2_blocked():
    """decode_strip_item (mirror of DCursor::decode): every (row block, chunk) with I < rpc (c + 1) exactly once for any
    super-row size; inside a super-row the items of one chunk are consecutive, and a super-row's row blocks are not
    touched again once it is finished (the packed rows of one super-row are what has to stay in the L2)."""
    from cglb_b200.distributed import decode_strip_item
    for n_chunks, rpc, q in [(1, 4, 1), (7, 4, 1), (7, 4, 2), (7, 4, 3), (7, 4, 100), (13, 2, 5), (40, 4, 8), (1953, 4, 512)]:
        nitems = rpc * n_chunks * (n_chunks + 1) // 2
        step = 1 if nitems < 200000 else 997           # the headline shape is sampled
        seen, last_sr = set(), -1
        for t in range(0, nitems, step):
            i, c = decode_strip_item(t, n_chunks, rpc, q)
            assert 0 <= c < n_chunks and 0 <= i < rpc * (c + 1), (t, i, c)
            assert (i, c) not in seen
            seen.add((i, c))
            sr = i // (rpc * q)
            assert sr >= last_sr                           # super-rows are visited in order, never revisited
            last_sr = sr
        if step == 1:
            assert len(seen) == nitems


def test_fixed_order_accumulation_has_one_owner_thread_per_address():
    """in the DMMA sweeps, address X of a CTA's private copy of y is added to by thread X % 256 only -- as a row
    (row r0 + t belongs to thread t) and as a column (column c0 + col is flushed by thread col % 256) -- because row blocks and
    the first column of every item are multiples of 256; program order of that one thread then fixes the summation order."""
    from cglb_b200.distributed import strip_items
    for n, q in ((5000, 1), (5000, 2), (70000, 16), (3000, 1 << 30)):
        for t, r0, r1, tiles in strip_items(n, 256, 4, 64, superrow_chunks=q):
            c0 = tiles[0][0]
            assert r0 % 256 == 0 and c0 % 256 == 0, (t, r0, c0)
            assert all(j0 % 64 == 0 for j0, _, _ in tiles)
            # columns that receive column sums (tiles beyond the row block) never overlap the item's own rows
            assert all(j0 >= r0 + 256 for j0, _, off in tiles if off)
