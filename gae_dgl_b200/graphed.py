"""CUDA-graph capture of a whole transductive training step (static graph, static shapes).

On citation-graph sizes a step is ~30 short kernels; launch and Python overhead is comparable to
the kernel time, so the loop is captured once and replayed (the "capture launch-bound inner
loops in CUDA graphs" rule).  Everything in the step is capture-safe: our kernels are launched on
torch's current (capturing) stream, workspaces come from torch's graph-private pool, the dropout
mask is drawn from a device-resident Philox state that advances on-stream (a fresh mask per
replay), and Adam runs with capturable=True.
"""
from __future__ import annotations

from typing import Callable

import torch


class GraphedTrainStep:
    """step() -> loss tensor (device, overwritten by every replay).

    loss_fn() must build the loss from static input tensors (update them in place between
    replays, e.g. `features.copy_(new)`), touching no host-side randomness."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable[[], torch.Tensor],
                 warmup: int = 3):
        for group in optimizer.param_groups:
            if not group.get("capturable", False):
                raise ValueError("GraphedTrainStep needs an optimizer built with capturable=True")
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.warmup_losses = []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):              # real optimisation steps, run eagerly
                optimizer.zero_grad(set_to_none=True)
                loss = loss_fn()
                loss.backward()
                optimizer.step()
                self.warmup_losses.append(loss.detach().clone())
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = loss_fn()
            self.loss.backward()
            optimizer.step()

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss
