"""Minimal stand-ins for the reference's observability helpers (cglb/backend/callbacks.py:27-178): a
pausable stop-watch and an in-memory logger with the same call protocol as the reference `Logger`
(`logger(step, variables, values)`, `log_for_feval(**kw)`, `no_recording()`, `.timer`).  TensorBoard output is
out of scope."""
from __future__ import annotations

import contextlib
import time
from typing import Callable, Dict, List, Optional


class StopWatch:
    def __init__(self):
        self.reset()

    def reset(self):
        self._elapsed, self._t0 = 0.0, None

    def start(self):
        if self._t0 is None:
            self._t0 = time.perf_counter()

    def stop(self):
        if self._t0 is not None:
            self._elapsed += time.perf_counter() - self._t0
            self._t0 = None

    @property
    def elapsed(self) -> float:
        return self._elapsed + (time.perf_counter() - self._t0 if self._t0 is not None else 0.0)

    @contextlib.contextmanager
    def pause(self):
        running = self._t0 is not None
        self.stop()
        try:
            yield
        finally:
            if running:
                self.start()


class Logger:
    def __init__(self, metrics_fn: Optional[Callable[[], Dict[str, float]]] = None, holdout_interval: int = 20):
        self.metrics_fn = metrics_fn
        self.holdout_interval = holdout_interval
        self.timer = StopWatch()
        self.logs: Dict[str, List] = {}
        self.feval_logs: Dict[str, List] = {}
        self._recording = True

    def log_for_feval(self, **kwargs):
        if not self._recording:
            return
        for k, v in kwargs.items():
            self.feval_logs.setdefault(k, []).append(float(v))

    @contextlib.contextmanager
    def no_recording(self):
        old, self._recording = self._recording, False
        try:
            yield
        finally:
            self._recording = old

    def __call__(self, step, *args):
        if not self._recording or self.holdout_interval <= 0 or step % self.holdout_interval != 0:
            return
        with self.timer.pause():
            record = {"step": step, "time": self.timer.elapsed}
            if self.metrics_fn is not None:
                record.update({k: float(v) for k, v in self.metrics_fn().items()})
            for k, v in record.items():
                self.logs.setdefault(k, []).append(v)
