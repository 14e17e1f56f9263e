"""The two global settings of the reference path: default float type (torch default dtype,
interface.py:94-116) and Cholesky jitter (gpytorch.settings.cholesky_jitter, interface.py:90-91,
read at models.py:192; 1e-6 for fp64 / 1e-5 for fp32, backend.py:76-79)."""


class _Setting:
    def __init__(self, value):
        self._value = value

    def value(self):
        return self._value

    def _set_value(self, value):
        self._value = value


cholesky_jitter = _Setting(1e-6)
