"""Backend facade + registry, as cglb/backend/backend.py:34-115: a `Backend` subclass whose `interface()` is
this package's interface module, registered under "b200" (and as a drop-in under "torch")."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, Tuple

import numpy as np

from . import interface as _interface

Data = Tuple[np.ndarray, np.ndarray]
Dataset = Tuple[Data, Data]

__all__ = ["Backend", "B200", "BACKENDS"]


class Backend(ABC):
    @staticmethod
    @abstractmethod
    def interface():
        pass

    @classmethod
    def configure_backend(cls, **kwargs):
        return cls.interface().configure_backend(**kwargs)

    @classmethod
    def create_kernel(cls, cfg, data: Data):
        return cls.interface().create_kernel(cfg, data)

    @classmethod
    def create_model(cls, model_cfg, data: Data):
        return cls.interface().create_model(model_cfg, data)

    @classmethod
    def model_parameters(cls, model) -> Dict[str, np.ndarray]:
        return cls.interface().model_parameters(model)

    @classmethod
    def optimize(cls, model, dataset: Dataset, num_steps: int, logger, optimizer: str):
        return cls.interface().optimize(model, dataset, num_steps, logger, optimizer)

    @classmethod
    def save(cls, model, logdir: str):
        return cls.interface().save(model, logdir)

    @classmethod
    def load(cls, model, filepath: str):
        return cls.interface().load(model, filepath)

    @classmethod
    def metrics_fn(cls, model, dataset_bundle: Tuple[Data, Data]):
        return cls.interface().metrics_fn(model, dataset_bundle)

    @classmethod
    def set_default_float(cls, float_type: str):
        return cls.interface().set_default_float(float_type)

    @classmethod
    def set_default_jitter(cls, float_type: str):
        value = 1e-5 if float_type == "fp32" else 1e-6            # backend.py:76-79
        return cls.interface().set_default_jitter(value)

    @classmethod
    def get_default_float_str(cls):
        return cls.interface().get_default_float_str()

    @classmethod
    def get_default_float(cls):
        return cls.interface().get_default_float()


class B200(Backend):
    @staticmethod
    def interface():
        return _interface


BACKENDS = {"b200": B200, "torch": B200}
