"""CPU: the hand-derived backward equals autograd through the oracle."""
import numpy as np
import pytest
import torch

from oracle import cglb_oracle as o
from closed_form_reference import closed_form_grads


@pytest.mark.parametrize("kind,n,d,M,noise", [("matern32", 150, 2, 12, 0.05), ("rbf", 120, 3, 10, 0.3),
                                              ("matern32", 90, 1, 8, 1.0)])
def test_closed_form_matches_autograd(kind, n, d, M, noise):
    x, y, z = o.synthetic_problem(n, d, M, seed=21)
    ls = torch.linspace(0.7, 1.3, d, dtype=torch.float64)
    p = o.OracleParams.from_values(noise, 0.07, z, 1.4, ls)
    v0 = torch.zeros(n, 1, dtype=torch.float64)
    loss, grads, res = o.bound_and_grads(kind, p, x, y, v0)
    # chain rule back from raw parameters: d/d(value) = d/d(raw) / sigmoid(raw)
    sig = torch.sigmoid
    g = closed_form_grads(kind, x, y, z, p.lengthscale.detach(), float(p.variance), float(p.noise), 0.07, res.v)
    # loss = -bound
    raw_noise_grad = -g["noise"] * float(sig(p.raw_noise))
    assert abs(raw_noise_grad - float(grads[0])) <= 1e-8 * abs(float(grads[0])) + 1e-10
    assert abs(-g["c"] - float(grads[1])) <= 1e-8 * abs(float(grads[1])) + 1e-9
    assert np.abs(-g["Z"].numpy() - grads[2].numpy()).max() <= 1e-8 * np.abs(grads[2].numpy()).max()
    raw_var_grad = -g["var"] * float(sig(p.raw_outputscale))
    assert abs(raw_var_grad - float(grads[3])) <= 1e-8 * abs(float(grads[3])) + 1e-10
    raw_ls_grad = -(g["ls"] * sig(p.raw_lengthscale.detach()).reshape(-1)).numpy()
    assert np.abs(raw_ls_grad - grads[4].numpy().reshape(-1)).max() <= 1e-8 * np.abs(grads[4].numpy()).max()
