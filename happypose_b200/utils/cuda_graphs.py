"""CUDA-graph replay of launch-bound inner loops.

One refiner step is ~170 small launches (7 libhpb200 kernels + the ResNet's ~40 cuDNN calls + heads) and a MegaPose pose
needs 5 of them back to back on ONE hypothesis, plus a one-row scoring pass: issued eagerly this is bound by the host's
launch rate (~10 us per torch op), not by the GPU.  GraphCache captures such a function once per (key, input shapes) on
torch's capture stream -- libhpb200 launches on torch's current stream, so its kernels are captured like torch's own --
and afterwards replays it with one cudaGraphLaunch.  Inputs are copied into the graph's static input tensors before
every replay; the function's outputs are the graph's static output tensors and are overwritten by the next replay, so
callers copy what they keep.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Dict, Hashable, Sequence, Tuple

import torch


class GraphCache:
    def __init__(self, epoch_fn: Callable[[], int] = None) -> None:
        # epoch_fn: the library context's workspace epoch.  libhpb200 never frees or moves a workspace (replaced buffers
        # are retired), so graphs captured before a growth stay valid; they are dropped anyway when the epoch moves so
        # that replays use the current buffers and the retired ones stop being touched.
        self._epoch_fn = epoch_fn
        self._epoch = None
        self._entries: Dict[Hashable, Tuple[torch.cuda.CUDAGraph, Tuple[torch.Tensor, ...], Any]] = {}
        self._failed: set = set()
        self.replays = 0
        self.captures = 0
        self.last_error = None

    @staticmethod
    def signature(tensors: Sequence[torch.Tensor]) -> Tuple:
        return tuple((tuple(t.shape), t.dtype, t.device.index) for t in tensors)

    def run(self, key: Hashable, fn: Callable[..., Any], tensors: Sequence[torch.Tensor]) -> Tuple[Any, bool]:
        """-> (outputs, replayed).  `fn(*tensors)` must be a pure function of its tensor arguments that launches only on
        the current stream and never synchronises the host.  When capture is impossible the function runs eagerly."""
        full_key = (key, self.signature(tensors))
        if self._epoch_fn is not None:
            epoch = self._epoch_fn()
            if self._epoch is not None and epoch != self._epoch:
                self._entries.clear()
            self._epoch = epoch
        if full_key in self._failed:
            return fn(*tensors), False
        entry = self._entries.get(full_key)
        if entry is None:
            static = tuple(t.clone() for t in tensors)
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):  # warm-up: cuDNN autotuning, libhpb200 workspace growth
                    for _ in range(2):
                        fn(*static)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = fn(*static)
            except Exception as exc:  # capture is an optimisation: fall back to eager execution, remember why
                torch.cuda.synchronize()
                self._failed.add(full_key)
                self.last_error = repr(exc)
                if os.environ.get("HPB_GRAPH_DEBUG"):
                    import traceback

                    traceback.print_exc()
                return fn(*tensors), False
            entry = (graph, static, out)
            self._entries[full_key] = entry
            self.captures += 1
            if self._epoch_fn is not None:
                self._epoch = self._epoch_fn()  # the warm-up itself may have grown a workspace
        graph, static, out = entry
        for s, t in zip(static, tensors):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        graph.replay()
        self.replays += 1
        return out, True

    def clear(self) -> None:
        self._entries.clear()
        self._failed.clear()
