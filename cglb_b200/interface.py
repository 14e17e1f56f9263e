"""Backend interface module: the functions the reference's `Backend` facade dispatches to
(cglb/backend/backend.py:34-91 -> cglb/backend/pytorch/interface.py), for the CGLB model only.

Same function names, argument meaning and error behaviour: `NotImplementedError` for unknown float
types / configs (interface.py:104,122), `AssertionError` if the optimiser is not "scipy" (interface.py:447).
Unlike the reference module this one does not import TensorFlow, pykeops or gpytorch.
"""
from __future__ import annotations

import json
import os
from dataclasses import asdict
from functools import singledispatch
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import settings
from ._ffi import CglbError
from .callbacks import Logger
from .config import CGLBConfig, KernelConfig, Matern32Config, ModelConfig, SGPRConfig, SquaredExponentialConfig
from .gp import GaussianLikelihood, GreaterThan, InducingPointKernel, MaternKernel, RBFKernel, ScaleKernel
from .models import CGLB, GPR, SGPR, LowerBoundCG, PredictCG, log_density
from .optimizer import Scipy

__all__ = ["configure_backend", "set_default_float", "set_default_jitter", "get_default_float", "get_default_float_str",
           "create_kernel", "create_model", "model_parameters", "optimize", "save", "load", "metrics_fn"]

Tensor = torch.Tensor
Data = Tuple[np.ndarray, np.ndarray]


def configure_backend(logdir: Optional[str] = None, keops: Optional[bool] = None, **kwargs):
    """interface.py:66-87.  `keops` is accepted for signature compatibility: the matrix-free sweep is always
    used (there is no dense n x n path and no JIT cache directory)."""
    assert logdir is not None
    assert keops is not None
    if not torch.cuda.is_available():
        raise CglbError("cglb_b200 needs a CUDA device (B200); there is no CPU fallback")


def set_default_jitter(jitter):                               # interface.py:90-91
    settings.cholesky_jitter._set_value(jitter)


def set_default_float(float_type: str) -> None:              # interface.py:94-104
    types = {"fp32": torch.float32, "float32": torch.float32, "fp64": torch.float64, "float64": torch.float64}
    if float_type in types:
        torch.set_default_dtype(types[float_type])
    else:
        raise NotImplementedError(f"Unknown float type {float_type}")


def get_default_float_str() -> str:
    return {torch.float32: "fp32", torch.float64: "fp64"}[torch.get_default_dtype()]


def get_default_float() -> np.dtype:
    return torch.tensor(1, dtype=torch.get_default_dtype()).detach().cpu().numpy().dtype


@singledispatch
def create_model(model_cfg: ModelConfig, data: Data):
    raise NotImplementedError()


@singledispatch
def create_kernel(cfg: KernelConfig, data: Data):
    raise NotImplementedError()


@singledispatch
def optimize(model: GPR, dataset: Tuple[Data, Data], num_steps: int, logdir: str, optimizer: str):
    raise NotImplementedError()


@singledispatch
def save(model: GPR, logdir: str):
    raise NotImplementedError()


@singledispatch
def load(model: GPR, filepath: str):
    raise NotImplementedError()


@singledispatch
def metrics_fn(model: GPR, dataset_bundle: Tuple[Data, Data]):
    raise NotImplementedError()


def model_parameters(model) -> Dict[str, np.ndarray]:       # interface.py:150-178
    noise = _numpy(model.likelihood.noise_covar.noise)[0]
    constant = _numpy(model.mean_module.constant)
    params = {".likelihood.variance": noise, ".mean_function.c": constant}
    kernel = model.covar_module
    if isinstance(kernel, InducingPointKernel):
        params.update({".inducing_variable.Z": _numpy(kernel.inducing_points)})
        kernel = kernel.base_kernel
    lengthscale = _numpy(kernel.base_kernel.lengthscale)[0, :]
    outputscale = _numpy(kernel.outputscale)
    params.update({".kernel.lengthscales": lengthscale, ".kernel.variance": outputscale.squeeze()})
    return params


# ---- implementations -------------------------------------------------------------------------------------
def _output_device() -> torch.device:
    if not torch.cuda.is_available():
        raise CglbError("cglb_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_tensor(array) -> Tensor:
    return torch.as_tensor(array, dtype=torch.get_default_dtype()).to(_output_device())


def _dataset_to_tensor(data: Tuple):
    return _to_tensor(data[0]), _to_tensor(data[1])


@create_kernel.register
def _create_kernel_rbf(cfg: SquaredExponentialConfig, data: Data):          # interface.py:207-217
    params = cfg.params(data)
    lengthscales = _to_tensor(params["lengthscales"])
    rbf = RBFKernel(ard_num_dims=len(lengthscales))
    rbf.lengthscale = lengthscales
    kernel = ScaleKernel(rbf)
    kernel.outputscale = _to_tensor(params["variance"])
    return kernel


@create_kernel.register
def _create_kernel_matern(cfg: Matern32Config, data: Data):                 # interface.py:220-230
    params = cfg.params(data)
    lengthscales = _to_tensor(params["lengthscales"])
    matern = MaternKernel(nu=1.5, ard_num_dims=len(lengthscales))
    matern.lengthscale = lengthscales
    kernel = ScaleKernel(matern)
    kernel.outputscale = _to_tensor(params["variance"])
    return kernel


def _likelihood_and_kernel_for_sgpr(model_cfg: SGPRConfig, data: Data):     # interface.py:263-301
    params = model_cfg.params(data)
    device = _output_device()
    likelihood = GaussianLikelihood(noise_constraint=GreaterThan(1e-6)).to(device)
    likelihood.noise = params["noise_variance"]
    base_kernel = create_kernel(model_cfg.kernel, data).to(device)

    def init_kernel_fn(x1, x2, full_cov: bool = False):
        x1 = _to_tensor(x1)
        if not full_cov:
            return base_kernel(x1, diag=True).detach().cpu().numpy()
        x2 = x1 if x2 is None else _to_tensor(x2)
        return base_kernel(x1, x2).evaluate().detach().cpu().numpy()

    init_kernel_fn.gpu_kernel = base_kernel      # lets InducingVariableConfig.init run the selection on the device
    inducing_variable = _to_tensor(params["inducing_variable"](init_kernel_fn))
    kernel = InducingPointKernel(base_kernel, inducing_variable, likelihood=likelihood)
    return likelihood, kernel


@create_model.register
def _create_model_cglb(model_cfg: CGLBConfig, data: Data):                  # interface.py:315-323
    likelihood, kernel = _likelihood_and_kernel_for_sgpr(model_cfg, data)
    data_tensors = (_to_tensor(data[0]), _to_tensor(np.asarray(data[1]).reshape(-1)))
    return CGLB(data_tensors, likelihood, kernel).to(_output_device())


@optimize.register
def _optimize_cglb(model: CGLB, dataset: Tuple[Data, Data], num_steps: int, logger: Logger, optimize: str):
    """Four SciPy L-BFGS-B phases, the last two without the inducing points (interface.py:445-543)."""
    assert optimize == "scipy"
    train_data = dataset[0]
    train_x = _to_tensor(train_data[0]).contiguous()
    train_y = _to_tensor(np.asarray(train_data[1]).reshape(-1)).contiguous()
    train_data = (train_x, train_y)
    model.train()
    lbfgs = Scipy()
    lower_bound = LowerBoundCG(model)

    def lbfgs_closure() -> Tensor:
        loss = -lower_bound(train_data)
        logger.log_for_feval(**asdict(model.cg_stats))
        return loss

    def step_callback(*args):
        lower_bound.cached_v_vec = False
        logger(*args)

    def optimize_fn(params, maxiter: int, ftol: float = 0.0, gtol: float = 0.0, disp: bool = False):
        options = dict(maxiter=maxiter, ftol=ftol, gtol=gtol, disp=disp)
        return lbfgs.minimize(lbfgs_closure, params, options=options, step_callback=step_callback)

    params = list(model.parameters())
    with logger.no_recording():                              # warm-up evaluation, excluded from timing (:495-501)
        _loss = lbfgs_closure()
        _grads = torch.autograd.grad(_loss, params)
        torch.cuda.synchronize()
    logger.timer.reset()
    logger.timer.start()

    results = []
    remaining = num_steps
    ips = model.covar_module.inducing_points
    for phase in range(4):
        if remaining <= 0:
            break
        if phase == 2:                                      # phases 3-4 drop Z from the variables (:527-529)
            params = [p for p in model.parameters() if id(p) != id(ips)]
        result = optimize_fn(params, remaining)
        remaining -= result.nit
        results.append(result)
    return results


@save.register
def _save(model: GPR, logdir: str):                          # interface.py:546-551 (json instead of json_tricks)
    os.makedirs(logdir, exist_ok=True)
    params = {k: np.asarray(v).tolist() for k, v in model_parameters(model).items()}
    with open(Path(logdir, "model.json"), "w") as file:
        json.dump(params, file)


@load.register
def _load(model: GPR, filepath: str):
    """Reads the JSON written by `save` (the reference's torch `load` expects a state-dict and is
    asymmetric with its `save`; the TF backend's JSON form is the one kept)."""
    with open(filepath) as f:
        params = json.load(f)
    model.likelihood.noise = torch.as_tensor(params[".likelihood.variance"])
    model.mean_module.constant.data.fill_(float(np.asarray(params[".mean_function.c"]).reshape(-1)[0]))
    kernel = model.covar_module
    if isinstance(kernel, InducingPointKernel):
        z = torch.as_tensor(params[".inducing_variable.Z"], dtype=kernel.inducing_points.dtype)
        kernel.inducing_points.data.copy_(z.to(kernel.inducing_points.device))
        kernel = kernel.base_kernel
    kernel.base_kernel.lengthscale = torch.as_tensor(params[".kernel.lengthscales"])
    kernel.outputscale = torch.as_tensor(params[".kernel.variance"])
    return model


@metrics_fn.register
def _compute_metrics_cglb(model: CGLB, dataset_bundle: Tuple[Data, Data]):  # interface.py:607-658
    def cglb_cg_params():
        if model.cg_stats is not None:
            return {"cg/steps": _numpy(model.cg_stats.steps), "cg/error": _numpy(model.cg_stats.residual_error)}
        return {}

    train, test = dataset_bundle
    data = _dataset_to_tensor((train[0], np.asarray(train[1]).reshape(-1)))

    def cglb_metrics():
        with torch.no_grad():
            lower_bound = LowerBoundCG(model, use_cache=True, cached_v_vec_initial=True)
            loss = -lower_bound(data)
            return dict(loss=_numpy(loss))

    x_full = torch.cat([_to_tensor(train[0]), _to_tensor(test[0])], 0)
    y_full = torch.cat([_to_tensor(train[1]).reshape(-1, 1), _to_tensor(test[1]).reshape(-1, 1)], 0)
    n_train = train[0].shape[0]

    def rmse_and_nlpd():
        predict_f = PredictCG(model)
        errs, lpds = [], []
        max_batch = int(1e6)
        with torch.no_grad():
            for i in range(0, x_full.shape[0], max_batch):
                f_mean, f_var = predict_f(x_full[i:i + max_batch])
                y_batch = y_full[i:i + max_batch]
                lpds.append(_numpy(log_density(model, y_batch, f_mean, f_var)))
                errs.append(_numpy(y_batch - f_mean))
        err, lpd = np.concatenate(errs, 0), np.concatenate(lpds, 0)
        out = {}
        for name, sl in (("train", slice(0, n_train)), ("test", slice(n_train, None))):
            if err[sl].size:
                out[f"{name}/rmse"] = float(np.sqrt(np.mean(err[sl] ** 2)))
                out[f"{name}/nlpd"] = float(-np.mean(lpd[sl]))
        return out

    def call():
        training = model.training
        model.eval()
        metrics = {}
        for cb in (cglb_cg_params, cglb_metrics, rmse_and_nlpd):
            metrics.update(cb())
        model.train(training)
        return metrics

    return call


def _numpy(tensor) -> np.ndarray:
    if isinstance(tensor, torch.Tensor):
        return tensor.detach().cpu().numpy()
    return np.asarray(tensor)
