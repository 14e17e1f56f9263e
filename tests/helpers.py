"""Shared test helpers: build a cglb_b200 model the way interface.py:263-323 does, from explicit values."""
import numpy as np
import torch

import cglb_b200 as cb


def make_model(kind, x, y, z, noise, variance, lengthscale, mean_c=0.0, device="cuda", dtype=torch.float64):
    d = x.shape[1]
    dev = torch.device(device)
    lik = cb.GaussianLikelihood(noise_constraint=cb.GreaterThan(1e-6)).to(dtype)
    lik.noise = noise
    base = (cb.MaternKernel(nu=1.5, ard_num_dims=d) if kind == "matern32" else cb.RBFKernel(ard_num_dims=d)).to(dtype)
    ls = torch.as_tensor(np.broadcast_to(np.asarray(lengthscale, dtype=np.float64).reshape(-1), (d,)).copy())
    base.lengthscale = ls
    scale = cb.ScaleKernel(base).to(dtype)
    scale.outputscale = variance
    ipk = cb.InducingPointKernel(scale, torch.as_tensor(z, dtype=dtype), likelihood=lik)
    xt = torch.as_tensor(x, dtype=dtype)
    yt = torch.as_tensor(y, dtype=dtype).reshape(-1)
    model = cb.CGLB((xt, yt), lik, ipk).to(dtype)
    model.mean_module.constant.data.fill_(mean_c)
    return model.to(dev)


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))
