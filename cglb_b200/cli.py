"""CLI shim for the reference's README command (cglb_experiments/cli.py:60-76,141-152,259-273,304-322):

    python -m cglb_b200.cli --keops -b b200 -t fp64 -l LOGDIR -s SEED train -n STEPS -d DATASET \
        cglb -k Matern32 -i ConditionalVariance -M 1024

Same option names and the same outputs (`model.json`, `results.json`, `logs.json` in LOGDIR,
cli.py:100-109).  The UCI/Wilson dataset loaders of the reference need the network and un-vendored
packages (datasets.py:47-76); here DATASET is `synthetic:<n>x<d>` or one of the BASELINE shapes
(`snelson1d`, `kin40k`, `3droad`, `song`, `houseelectric`), generated as in SURVEY.md section 8d and
z-score normalised like datasets.py:35-39.  `--seed` selects the split, as in the reference.
"""
from __future__ import annotations

import json
import math
import tempfile
from pathlib import Path

import click
import numpy as np

from .backend import BACKENDS
from .callbacks import Logger
from .config import CGLBConfig, INDUCING_VARIABLE_CONFIGS, KERNEL_CONFIGS

_SHAPES = {"snelson1d": (2000, 1), "kin40k": (40000, 8), "3droad": (434000, 3), "song": (515000, 90),
           "houseelectric": (2000000, 11)}


def make_dataset(name: str, seed: int, dtype):
    if name.startswith("synthetic:"):
        n, d = (int(t) for t in name.split(":", 1)[1].lower().split("x"))
    elif name in _SHAPES:
        n, d = _SHAPES[name]
    else:
        raise click.BadParameter(f"unknown dataset {name!r}: use synthetic:<n>x<d> or one of {sorted(_SHAPES)}")
    rng = np.random.RandomState(seed)
    n_test = max(1, n // 10)
    x = rng.randn(n + n_test, d)
    w, w2 = rng.randn(d), rng.randn(d)
    f = np.sin(2.0 * (x @ w) / math.sqrt(d)) + 0.5 * np.cos((x @ w2) / math.sqrt(d))
    y = (f + 0.1 * rng.randn(n + n_test)).reshape(-1, 1)
    xm, xs, ym, ys = x[:n].mean(0), x[:n].std(0), y[:n].mean(0), y[:n].std(0)       # datasets.py:35-39
    x, y = (x - xm) / xs, (y - ym) / ys
    return (x[:n].astype(dtype), y[:n].astype(dtype)), (x[n:].astype(dtype), y[n:].astype(dtype))


@click.group()
@click.option("-b", "--backend", type=click.Choice(sorted(BACKENDS)), required=True)
@click.option("-t", "--float-type", type=click.Choice(["fp32", "fp64"]), default="fp32")
@click.option("-l", "--logdir", type=click.Path(file_okay=False), default=None)
@click.option("-s", "--seed", type=int, default=0)
@click.option("--keops/--no-keops", default=True)
@click.pass_context
def main(ctx, backend, float_type, logdir, seed, keops):
    logdir = str(Path(logdir or tempfile.mkdtemp(prefix="cglb_b200_")).expanduser().resolve())
    Path(logdir).mkdir(exist_ok=True, parents=True)
    be = BACKENDS[backend]
    be.configure_backend(logdir=logdir, keops=keops)
    be.set_default_float(float_type)
    be.set_default_jitter(float_type)
    ctx.obj = dict(backend=be, seed=seed, logdir=logdir)


@main.group()
@click.option("-n", "--num-steps", type=int, default=1000)
@click.option("-d", "--dataset", type=str, required=True)
@click.option("-o", "--optimizer", type=click.Choice(["scipy"]), default="scipy")
@click.pass_context
def train(ctx, num_steps, dataset, optimizer):
    be = ctx.obj["backend"]
    ctx.obj.update(num_steps=num_steps, optimizer=optimizer,
                   dataset=make_dataset(dataset, ctx.obj["seed"], be.get_default_float()))


@train.command()
@click.option("-k", "--kernel", type=click.Choice(sorted(KERNEL_CONFIGS)), default="Matern32")
@click.option("-m", "--model", "model_name", type=str, default="cglb")
@click.option("-i", "--inducing-variable", type=click.Choice(sorted(INDUCING_VARIABLE_CONFIGS)), default="ConditionalVariance")
@click.option("-M", "--num-inducing-variables", type=int, default=1024)
@click.option("-e", "--max-error", type=float, default=1.0)
@click.pass_context
def cglb(ctx, kernel, model_name, inducing_variable, num_inducing_variables, max_error):
    o = ctx.obj
    be, (train_data, test_data) = o["backend"], o["dataset"]
    cfg = CGLBConfig(kernel=KERNEL_CONFIGS[kernel](), inducing_variable=INDUCING_VARIABLE_CONFIGS[inducing_variable](num_inducing_variables),
                     max_error=max_error)
    model = be.create_model(cfg, train_data)
    metrics_fn = be.metrics_fn(model, (train_data, test_data))
    logger = Logger(metrics_fn, holdout_interval=20)
    be.optimize(model, (train_data, test_data), o["num_steps"], logger, o["optimizer"])       # cli.py:87-109
    be.save(model, o["logdir"])
    results = {k: float(v) for k, v in metrics_fn().items()}
    results["id"] = o["logdir"]
    logs = dict(logger.logs, feval=logger.feval_logs, id=o["logdir"])
    with open(Path(o["logdir"], "results.json"), "w") as f:
        json.dump(results, f)
    with open(Path(o["logdir"], "logs.json"), "w") as f:
        json.dump(logs, f)
    click.echo(json.dumps(results))


if __name__ == "__main__":
    main()
