"""Timers with the reference's interfaces: SimpleTimer / CudaTimer (megapose/training/utils.py:218-277) and the
pause/resume Timer (toolbox/utils/timer.py:20-51).  CudaTimer records on torch's current stream and, like the
reference, synchronises in end(); with enabled=False it is a no-op that reports 0 s."""
from __future__ import annotations

import datetime
import time

import torch


class SimpleTimer:
    def __init__(self) -> None:
        self.start_time = None
        self.end_time = None

    def start(self):
        self.start_time = time.time()

    def stop(self):
        self.end_time = time.time()

    end = stop

    def elapsed(self) -> float:
        return self.end_time - self.start_time


class CudaTimer:
    def __init__(self, enabled: bool = True) -> None:
        self.enabled = enabled
        self._t0 = self._t1 = None
        self.elapsed_sec = None

    def start(self) -> None:
        if self.enabled:
            self._t0 = torch.cuda.Event(enable_timing=True)
            self._t1 = torch.cuda.Event(enable_timing=True)
            self._t0.record()

    def end(self) -> None:
        if not self.enabled:
            return
        if self._t0 is None:
            raise ValueError("You must call CudaTimer.start() before CudaTimer.end()")
        self._t1.record()
        self._t1.synchronize()
        self.elapsed_sec = self._t0.elapsed_time(self._t1) / 1000.0

    stop = end

    def elapsed(self) -> float:
        if not self.enabled:
            return 0.0
        if self.elapsed_sec is None:
            raise ValueError("You must call CudaTimer.start() and CudaTimer.end() before querying the elapsed time")
        return self.elapsed_sec


class Timer:
    def __init__(self):
        self.reset()
        self.elapsed = datetime.timedelta()

    def reset(self):
        self.start_time = None
        self.elapsed = 0.0
        self.is_running = False

    def start(self):
        self.elapsed = datetime.timedelta()
        self.is_running = True
        self.start_time = datetime.datetime.now()
        return self

    def pause(self):
        if self.is_running:
            self.elapsed += datetime.datetime.now() - self.start_time
            self.is_running = False

    def resume(self):
        if not self.is_running:
            self.start_time = datetime.datetime.now()
            self.is_running = True

    def stop(self):
        self.pause()
        elapsed = self.elapsed
        self.reset()
        return elapsed
