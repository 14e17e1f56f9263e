"""Frozen-dataclass configs and name->class registries of the reference (cglb/backend/config.py:45-166),
restricted to what the CGLB path uses.  Initial hyper-parameters: lengthscales 1, variance 1, noise 1
(config.py:76,105)."""
from __future__ import annotations

import dataclasses
from functools import partial
from typing import Callable, Dict, Tuple, Union

import numpy as np

__all__ = ["Config", "ModelConfig", "KernelConfig", "SquaredExponentialConfig", "Matern32Config", "SGPRConfig",
           "CGLBConfig", "InducingVariableConfig", "SGPR_CONFIGS", "KERNEL_CONFIGS", "INDUCING_VARIABLE_CONFIGS"]

Data = Tuple[np.ndarray, np.ndarray]
dataclass_frozen = partial(dataclasses.dataclass, frozen=True)


class Config:
    def params(self, **kwargs) -> Dict[str, Union[float, np.ndarray]]:
        pass


@dataclass_frozen
class ModelConfig(Config):
    pass


@dataclass_frozen
class InducingVariableConfig(Config):
    num_variables: int

    def params(self, data: Data) -> Dict[str, Union[float, np.ndarray]]:
        ...

    def init(self, data: Data, kernel_fn: Callable):
        """Greedy conditional-variance selection (robustgp.ConditionalVariance(sample=False), config.py:62-65)."""
        from .inducing import ConditionalVariance, conditional_variance_gpu
        gpu_kernel = getattr(kernel_fn, "gpu_kernel", None)
        if gpu_kernel is not None:            # device-resident variant (interface.py passes the kernel module along)
            iv, _ = conditional_variance_gpu(data[0], self.num_variables, gpu_kernel)
        else:
            iv, _ = ConditionalVariance(sample=False)(data[0], self.num_variables, kernel_fn)
        return iv


class KernelConfig(Config):
    pass


@dataclass_frozen
class SquaredExponentialConfig(KernelConfig):
    def params(self, data: Data) -> Dict[str, Union[float, np.ndarray]]:
        vecdim = data[0].shape[-1]
        return {"variance": 1.0, "lengthscales": np.repeat(1.0, vecdim)}


@dataclass_frozen
class Matern32Config(SquaredExponentialConfig):
    pass


@dataclass_frozen
class SGPRConfig(ModelConfig):
    kernel: KernelConfig
    inducing_variable: InducingVariableConfig

    def params(self, data: Data) -> Dict[str, Union[float, np.ndarray, Callable]]:
        return {"noise_variance": 1.0, "inducing_variable": partial(self.inducing_variable.init, data)}


@dataclass_frozen
class CGLBConfig(SGPRConfig):
    max_error: float = 1.0
    joint_optimization: bool = False
    vzero: bool = False

    def params(self, data: Data) -> Dict[str, Union[float, np.ndarray]]:
        param_dict = super().params(data)
        param_dict["max_error"] = self.max_error
        param_dict["joint_optimization"] = self.joint_optimization
        param_dict["vzero"] = self.vzero
        return param_dict


SGPR_CONFIGS = {"cglb": CGLBConfig}
KERNEL_CONFIGS = {"SquaredExponential": SquaredExponentialConfig, "Matern32": Matern32Config,
                  "mat32": Matern32Config, "rbf": SquaredExponentialConfig}
INDUCING_VARIABLE_CONFIGS = {"InducingVariable": InducingVariableConfig, "ConditionalVariance": InducingVariableConfig,
                             "iv": InducingVariableConfig, "cv": InducingVariableConfig}
