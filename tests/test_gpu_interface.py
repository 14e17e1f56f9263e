"""GPU (-m gpu): the backend interface end to end -- create_model (ConditionalVariance init), the 4-phase
SciPy optimisation loop (interface.py:445-543), metrics, save/load -- on a small synthetic problem."""
import json
import os

import numpy as np
import pytest
import torch

import cglb_b200 as cb
from cglb_b200 import interface
from cglb_b200.callbacks import Logger
from oracle import cglb_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fp64_default():
    old = torch.get_default_dtype()
    interface.set_default_float("fp64")
    cb.B200.set_default_jitter("fp64")
    yield
    torch.set_default_dtype(old)


def _data(n=600, d=2, seed=3):
    x, y, _ = o.synthetic_problem(n + 100, d, 4, seed=seed)
    x, y = x.numpy(), y.numpy().reshape(-1, 1)
    return (x[:n], y[:n]), (x[n:], y[n:])


def test_create_optimize_metrics_save_load(fp64_default, tmp_path):
    train, test = _data()
    backend = cb.BACKENDS["b200"]
    backend.configure_backend(logdir=str(tmp_path), keops=True)
    assert backend.get_default_float_str() == "fp64" and backend.get_default_float() == np.float64
    cfg = cb.CGLBConfig(kernel=cb.Matern32Config(), inducing_variable=cb.InducingVariableConfig(32))
    model = backend.create_model(cfg, train)
    assert isinstance(model, cb.CGLB) and model.covar_module.inducing_points.shape == (32, 2)
    # initial hyper-parameters of the reference (config.py:76,105)
    pars = backend.model_parameters(model)
    assert abs(float(pars[".likelihood.variance"]) - 1.0) < 1e-9 and np.allclose(pars[".kernel.lengthscales"], 1.0)
    # inducing points are a subset of the training inputs (greedy conditional variance)
    z = pars[".inducing_variable.Z"]
    assert all(np.any(np.all(np.isclose(train[0], zi), axis=1)) for zi in z)

    metrics = backend.metrics_fn(model, (train, test))
    logger = Logger(metrics, holdout_interval=5)
    lb = cb.LowerBoundCG(model)
    loss0 = float(-lb((model.train_inputs[0], model.train_targets)))
    results = backend.optimize(model, (train, test), 12, logger, "scipy")
    assert sum(r.nit for r in results) >= 1
    loss1 = float(-lb((model.train_inputs[0], model.train_targets)))
    assert loss1 < loss0 - 1.0                                  # the bound improved
    assert len(logger.feval_logs["steps"]) >= 1                 # CG statistics were logged per f-eval
    m = metrics()
    assert {"cg/steps", "cg/error", "loss", "train/rmse", "test/rmse", "train/nlpd", "test/nlpd"} <= set(m)
    assert m["test/rmse"] < 1.0 and np.isfinite(m["test/nlpd"])
    with pytest.raises(AssertionError):
        backend.optimize(model, (train, test), 1, logger, "adam_0.1")

    backend.save(model, str(tmp_path))
    saved = json.load(open(os.path.join(tmp_path, "model.json")))
    assert set(saved) == {".likelihood.variance", ".mean_function.c", ".inducing_variable.Z", ".kernel.lengthscales", ".kernel.variance"}
    model2 = backend.create_model(cfg, train)
    backend.load(model2, os.path.join(tmp_path, "model.json"))
    p1, p2 = backend.model_parameters(model), backend.model_parameters(model2)
    for k in p1:
        assert np.allclose(p1[k], p2[k], rtol=1e-10, atol=1e-12), k


def test_fp32_models_use_fp32_pair_sweeps(fp64_default):
    """fp32 switch (interface.py:94-104): parameters and data may be fp32; the n^2 kernel-pair evaluations of the K*v
    sweeps then run in FP32 (cglb_kmv_sym_f32), everything else is promoted to fp64.  Tolerance 1e-4 relative on the
    bound (fp32 pair evaluations: ~1e-6 per entry)."""
    interface.set_default_float("fp32")
    try:
        train, _ = _data(n=300)
        cfg = cb.CGLBConfig(kernel=cb.SquaredExponentialConfig(), inducing_variable=cb.InducingVariableConfig(16))
        model = cb.B200.create_model(cfg, train)
        assert model.train_inputs[0].dtype == torch.float32
        loss = -cb.LowerBoundCG(model)((model.train_inputs[0], model.train_targets))
        grads = torch.autograd.grad(loss, list(model.parameters()))
        assert loss.dtype == torch.float32 and all(g.dtype == torch.float32 and torch.isfinite(g).all() for g in grads)
        x64, y64 = torch.as_tensor(train[0], dtype=torch.float64), torch.as_tensor(train[1], dtype=torch.float64).reshape(-1)
        z = model.covar_module.inducing_points.detach().double().cpu()
        p = o.OracleParams.from_values(1.0, 0.0, z, 1.0, 1.0)
        ref = o.lower_bound("rbf", p, x64.float().double(), y64.float().double(), torch.zeros(300, 1, dtype=torch.float64))
        assert abs(float(loss) + float(ref.bound)) <= 1e-4 * abs(float(ref.bound))
    finally:
        interface.set_default_float("fp64")


def test_lbfgs_trajectory_matches_the_oracle(fp64_default):
    """SURVEY.md 8f-4, end-to-end driver parity: the same L-BFGS-B run (optimizer.py:21-48: flat fp64 vector in, loss and
    packed gradients out, CG warm start carried across function evaluations, models.py:274) over the device path and over
    the CPU oracle, from the same starting point.  SciPy is deterministic, so the two runs take the same iterates as long
    as bound and gradients agree: same number of function evaluations, the loss after every evaluation <= 1e-7 apart
    (north_star), the final parameters <= 1e-5 apart (8 iterations amplify the 1e-8 gradient differences a little)."""
    import scipy.optimize
    n, d, M = 500, 2, 16
    x, y, z = o.synthetic_problem(n, d, M, seed=9)
    noise, var, ls = 0.3, 1.0, 1.0
    # ---- device path through the reference-shaped driver
    from helpers import make_model
    model = make_model("matern32", x.numpy(), y.numpy(), z.numpy(), noise, var, ls, 0.0)
    lb = cb.LowerBoundCG(model)
    data = (model.train_inputs[0], model.train_targets)
    gpu_losses = []

    def closure():
        loss = -lb(data)
        gpu_losses.append(float(loss))
        return loss

    params = list(model.parameters())
    x0 = cb.Scipy.to_numpy(cb.Scipy.pack(params)).astype(np.float64).copy()
    res_gpu = cb.Scipy().minimize(closure, params, options=dict(maxiter=8, ftol=0.0, gtol=0.0))
    # ---- the oracle through scipy directly (same packing order: raw noise, mean constant, Z, raw outputscale, raw lengthscale)
    p = o.OracleParams.from_values(noise, 0.0, z, var, ls)
    state = {"v": torch.zeros(n, 1, dtype=torch.float64)}
    cpu_losses = []

    def f(vec):
        t = torch.from_numpy(vec)
        off = 0
        for tens in p.tensors():
            cnt = tens.numel()
            tens.data = t[off:off + cnt].reshape(tens.shape).clone()
            off += cnt
        loss, grads, res = o.bound_and_grads("matern32", p, x, y, state["v"])
        state["v"] = res.v
        cpu_losses.append(float(loss))
        return float(loss), torch.cat([g.reshape(-1) for g in grads]).numpy().astype(np.float64)

    x0_cpu = torch.cat([t.detach().reshape(-1) for t in p.tensors()]).numpy().astype(np.float64)
    assert np.allclose(x0, x0_cpu, rtol=0, atol=1e-14)
    res_cpu = scipy.optimize.minimize(f, x0_cpu, jac=True, method="L-BFGS-B", options=dict(maxiter=8, ftol=0.0, gtol=0.0))
    assert res_gpu.nit == res_cpu.nit and res_gpu.nfev == res_cpu.nfev and len(gpu_losses) == len(cpu_losses)
    for a, b in zip(gpu_losses, cpu_losses):
        assert abs(a - b) <= 1e-7 * max(abs(b), 10.0)           # the loss crosses zero on the way (354 -> -86)
    assert np.abs(res_gpu.x - res_cpu.x).max() <= 1e-5 * np.abs(res_cpu.x).max()
    assert gpu_losses[-1] < gpu_losses[0] - 1.0


def test_fp32_models_predict_and_metrics(fp64_default, tmp_path):
    """an fp32 model must get through PredictCG, the sub-step methods and the Logger's metrics callback
    (interface.py:607-658) -- every operand is promoted to the fp64 kernels and the results come back as fp32."""
    interface.set_default_float("fp32")
    try:
        train, test = _data(n=400)
        cfg = cb.CGLBConfig(kernel=cb.Matern32Config(), inducing_variable=cb.InducingVariableConfig(16))
        model = cb.B200.create_model(cfg, train)
        assert model.train_inputs[0].dtype == torch.float32
        mean, var = cb.PredictCG(model)(torch.as_tensor(test[0], dtype=torch.float32).cuda())
        assert mean.dtype == torch.float32 and var.dtype == torch.float32 and mean.shape == (100, 1)
        assert torch.isfinite(mean).all() and (var > 0).all()
        # the same prediction from an fp64 copy of the model: fp32 kernel pairs cost ~1e-6 per entry
        interface.set_default_float("fp64")
        model64 = cb.B200.create_model(cfg, train)
        mean64, var64 = cb.PredictCG(model64)(torch.as_tensor(test[0], dtype=torch.float64).cuda())
        assert float((mean.double() - mean64).abs().max()) <= 1e-3 * float(mean64.abs().max()) + 1e-4
        assert float((var.double() - var64).abs().max()) <= 1e-3 * float(var64.abs().max())
        interface.set_default_float("fp32")
        lb = cb.LowerBoundCG(model)
        terms = lb.logdet_and_quad_common_terms((model.train_inputs[0], model.train_targets))
        assert terms.LB.shape == (16, 16)
        m = cb.B200.metrics_fn(model, (train, test))()
        assert np.isfinite(m["test/rmse"]) and np.isfinite(m["test/nlpd"]) and np.isfinite(m["loss"])
        # the solver API called directly with fp32 / strided tensors: promoted, not misread
        xd = model.train_inputs[0]
        kern = model.covar_module.base_kernel
        cov = kern(xd).add_diag(model.likelihood.noise.squeeze()).detach()
        b = torch.randn(400, 2, dtype=torch.float32, device=xd.device)[:, :1]          # strided fp32 view
        A = torch.randn(8, 400, dtype=torch.float32, device=xd.device) * 0.1
        LB = torch.linalg.cholesky(torch.eye(8, device=xd.device) + A @ A.t())
        pre = cb.NystromPreconditioner(A, LB, torch.tensor(1.0, device=xd.device))
        v, st = cb.ConjugateGradient()(cov, b, torch.zeros_like(b), pre)
        assert v.dtype == torch.float32 and torch.isfinite(v).all() and int(st.steps) >= 0
        from cglb_b200.engine import get_engine
        eng = get_engine()
        with pytest.raises(cb.CglbError):
            eng.dot(b.reshape(-1).contiguous(), b.reshape(-1).contiguous(), eng.empty(1))      # fp32 handed to an fp64 kernel
    finally:
        interface.set_default_float("fp64")


def test_readme_command_shim(tmp_path):
    """README command of the reference (README.md:35) through the CLI shim, on a small synthetic dataset."""
    from click.testing import CliRunner
    from cglb_b200.cli import main
    old = torch.get_default_dtype()
    try:
        res = CliRunner().invoke(main, ["--keops", "-b", "b200", "-t", "fp64", "-l", str(tmp_path), "-s", "1", "train", "-n", "6",
                                        "-d", "synthetic:500x2", "cglb", "-k", "Matern32", "-i", "ConditionalVariance", "-M", "24"],
                                 catch_exceptions=False)
        assert res.exit_code == 0, res.output
        for name in ("model.json", "results.json", "logs.json"):
            assert os.path.isfile(os.path.join(tmp_path, name))
        results = json.load(open(os.path.join(tmp_path, "results.json")))
        assert {"loss", "test/rmse", "test/nlpd", "cg/steps"} <= set(results) and results["test/rmse"] < 1.0
    finally:
        torch.set_default_dtype(old)


def test_gpu_conditional_variance_matches_host_version(fp64_default):
    from cglb_b200.inducing import ConditionalVariance, conditional_variance_gpu
    train, _ = _data(n=3000, d=3)
    kernel = interface.create_kernel(cb.Matern32Config(), train).cuda()
    kernel.base_kernel.lengthscale = torch.tensor([0.7, 1.0, 1.4], dtype=torch.float64)

    def host_kernel(x1, x2, full_cov=False):
        x1t = torch.as_tensor(x1, dtype=torch.float64).cuda()
        if not full_cov:
            return kernel(x1t, diag=True).detach().cpu().numpy()
        x2t = x1t if x2 is None else torch.as_tensor(x2, dtype=torch.float64).cuda()
        return kernel(x1t, x2t).evaluate().detach().cpu().numpy()

    z_cpu, idx_cpu = ConditionalVariance(sample=False)(train[0], 48, host_kernel)
    z_gpu, idx_gpu = conditional_variance_gpu(train[0], 48, kernel)
    assert np.array_equal(idx_cpu, idx_gpu) and np.array_equal(z_cpu, z_gpu)
    # Independent of both implementations: the DEFINITION of the published algorithm (robustgp is absent here, so this is
    # what "ConditionalVariance" can be pinned to) -- every selected point maximises the variance conditioned on the points
    # selected before it, var_i - k_iS K_SS^-1 k_Si, evaluated densely with the ORACLE's kernel arithmetic on the CPU.
    x = torch.as_tensor(train[0], dtype=torch.float64)
    ls, var = torch.tensor([0.7, 1.0, 1.4], dtype=torch.float64), torch.tensor(1.0, dtype=torch.float64)
    assert len(set(idx_gpu.tolist())) == 48
    chosen = [int(idx_gpu[0])]
    for step in range(1, 48):
        xs = x[chosen]
        kss = o.kernel_dense("matern32", xs, xs, ls, var) + 1e-12 * torch.eye(len(chosen), dtype=torch.float64)
        kxs = o.kernel_dense("matern32", x, xs, ls, var)
        cond = var - (kxs * torch.linalg.solve(kss, kxs.T).T).sum(1)
        assert float(cond[int(idx_gpu[step])]) >= float(cond.max()) - 1e-9, step
        chosen.append(int(idx_gpu[step]))
