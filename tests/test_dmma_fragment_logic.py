"""CPU: the index algebra of the DMMA sweeps (cglb_b200/csrc/dsweep_impl.cuh), emulated lane by lane in numpy.

mma.sync.m8n8k4 (f64): lane = 4 g + t4 holds A[g][t4], B[t4][g] and C[g][2 t4], C[g][2 t4 + 1].
  * forward: the A fragment carries -2 a_k and 1.0 in the slot of |b|^2, the first DMMA of a chain starts from
    (|a|^2, |a|^2): the chain ends with q = |a|^2 + |b|^2 - 2 a.b, also when the contraction is padded past the packed
    row into the next one (A carries 0 there);
  * backward: the thread's own two C values (columns 2 t4, 2 t4 + 1) are reused as A fragments of two k-steps whose B
    fragments are rows 2 t4 / 2 t4 + 1 of the column tile -- a permutation of the contraction index -- and the result is
    Y = C A_J."""
import numpy as np


def dmma_8x8x4(a_frag, b_frag, c_frag):
    """a_frag[lane], b_frag[lane]: one double per lane; c_frag[lane, 2].  Returns d_frag[lane, 2]."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        A[g, t4] = a_frag[lane]
        B[t4, g] = b_frag[lane]
        C[g, 2 * t4], C[g, 2 * t4 + 1] = c_frag[lane]
    D = A @ B + C
    return np.array([[D[lane >> 2, 2 * (lane & 3)], D[lane >> 2, 2 * (lane & 3) + 1]] for lane in range(32)])


def _packed(x):
    """cglb_pack_inputs layout: d coordinates, zero padding, |.|^2 in the last slot, width d + 1 rounded up to even."""
    n, d = x.shape
    dp = (d + 2) & ~1
    xp = np.zeros((n, dp))
    xp[:, :d] = x
    xp[:, dp - 1] = (x * x).sum(1)
    return xp, dp


def _forward_tile(xp_rows, xp_cols_flat, dp, col0):
    """q for rows 0..7 (one m-tile) x columns col0..col0+7, as dmma_ntile computes it."""
    ks_n = (dp + 3) // 4
    acc = None
    for ks in range(ks_n):
        a_frag = np.zeros(32); b_frag = np.zeros(32)
        for lane in range(32):
            g, t4 = lane >> 2, lane & 3
            k = 4 * ks + t4
            val = xp_rows[g, k] if k < dp else 0.0
            a_frag[lane] = 1.0 if k == dp - 1 else -2.0 * val
            b_frag[lane] = xp_cols_flat[(col0 + g) * dp + k]        # may run into the next packed row
        if acc is None:
            acc = np.array([[xp_rows[lane >> 2, dp - 1]] * 2 for lane in range(32)])      # C = (|a|^2, |a|^2)
        acc = dmma_8x8x4(a_frag, b_frag, acc)
    q = np.zeros((8, 8))
    for lane in range(32):
        g, t4 = lane >> 2, lane & 3
        q[g, 2 * t4], q[g, 2 * t4 + 1] = acc[lane]
    return q


def test_forward_chain_yields_squared_distances_for_every_width():
    rng = np.random.default_rng(0)
    for d in (2, 3, 8, 9, 11, 12, 13, 16, 27, 32):
        a = rng.standard_normal((8, d)); b = rng.standard_normal((17, d))
        ap, dp = _packed(a)
        bp, _ = _packed(b)
        flat = np.concatenate([bp.reshape(-1), rng.standard_normal(8)])      # finite data after the tile, as in shared memory
        for col0 in (0, 8):
            q = _forward_tile(ap, flat, dp, col0)
            ref = ((a[:, None, :] - b[None, col0:col0 + 8, :]) ** 2).sum(-1)
            assert np.allclose(q, ref, rtol=1e-12, atol=1e-12), d


def test_backward_cross_term_with_permuted_contraction():
    rng = np.random.default_rng(1)
    for d in (6, 8, 11, 13, 19):
        nq = (d + 7) // 8
        b = rng.standard_normal((9, d))
        bp, dp = _packed(b)
        flat = np.concatenate([bp.reshape(-1), rng.standard_normal(16)])
        cw = rng.standard_normal((8, 8))                    # c = e' omega for rows g, columns 0..7 (C-fragment layout)
        y = np.zeros((nq, 32, 2))
        for e in range(2):
            for q8 in range(nq):
                a_frag = np.array([cw[lane >> 2, 2 * (lane & 3) + e] for lane in range(32)])       # own C value as A[g][t4]
                b_frag = np.array([flat[(2 * (lane & 3) + e) * dp + 8 * q8 + (lane >> 2)] for lane in range(32)])
                y[q8] = dmma_8x8x4(a_frag, b_frag, y[q8])
        Y = np.zeros((8, 8 * nq))
        for q8 in range(nq):
            for lane in range(32):
                g, t4 = lane >> 2, lane & 3
                Y[g, 8 * q8 + 2 * t4], Y[g, 8 * q8 + 2 * t4 + 1] = y[q8, lane]
        assert np.allclose(Y[:, :d], cw @ b[:8], rtol=1e-12, atol=1e-12), d     # slots >= d are padding and ignored
