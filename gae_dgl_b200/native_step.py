"""One inductive train step without the autograd engine and without torch.optim's Python layer (the default of
train_inductive.py; `--no_native_step` restores loss.backward() + optim.step()).  Measured on the B200
(profiles/r02_round2_checks.log, ZINC-shaped batch = 256): 0.213 ms per step against 0.657 ms through autograd,
loss within 1.6e-7 relative and weights within 4.5e-8 of torch.optim.Adam after 6 steps
(tests/test_parity_gpu.py::test_native_train_step_follows_autograd_and_torch_adam).

`tools/zinc_profile.py` on the B200: a ZINC-shaped batch = 256 step takes 0.58 ms of host time for
about 0.25 ms of device work; what remains after the fused native step (`gae_step_fwd_bwd_f32`) is
torch's machinery around it -- `Function.apply` + `run_backward` (they only move already-computed
gradients into `.grad`) and `Optimizer.step` (state bookkeeping, a foreach add for the step counter,
the fused Adam launch).  Here `Trainer.iteration`'s three lines (train_inductive.py:50-52)

    optim.zero_grad(); loss.backward(); optim.step()

become two FFI calls: `gae_step_fwd_bwd_f32` writes dW / db straight into persistent `.grad` buffers,
`gae_adam_step_f32` applies torch.optim.Adam's update to every parameter in one launch.  The Adam
state lives in the wrapped `torch.optim.Adam` object under torch's own keys (`step`, `exp_avg`,
`exp_avg_sq`), so `optim.state_dict()` / `load_state_dict()` and the `.ckpt` files stay interchangeable
with the autograd path.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib, ops
from ._lib import GaeError, StepDesc
from .gae import GAE, _act_code, pos_weight_of


class NativeTrainStep:
    def __init__(self, model: GAE, optim: torch.optim.Adam):
        if not isinstance(model, GAE):
            raise GaeError("NativeTrainStep drives gae.GAE (the VGAE loss has an autograd-only KL term)")
        if len(optim.param_groups) != 1:
            raise GaeError("NativeTrainStep expects a single parameter group")
        grp = optim.param_groups[0]
        if grp.get("weight_decay", 0) or grp.get("amsgrad", False) or grp.get("maximize", False):
            raise GaeError("NativeTrainStep implements plain Adam (no weight decay / amsgrad / maximize)")
        self.model, self.optim, self.group = model, optim, grp
        self.codes = [_act_code(conv.apply_mod.activation) for conv in model.layers]
        if any(c is None for c in self.codes):
            raise GaeError("NativeTrainStep needs ReLU / identity activations")
        self.Ws = [conv.apply_mod.linear.weight for conv in model.layers]
        self.bs = [conv.apply_mod.linear.bias for conv in model.layers]
        self.params = [t for pair in zip(self.Ws, self.bs) for t in pair]
        if len(self.params) > 16 or any(not p.is_cuda or not p.is_contiguous() for p in self.params):
            raise GaeError("NativeTrainStep needs <= 16 contiguous CUDA parameter tensors")
        if {id(p) for p in self.params} != {id(p) for p in grp["params"]}:
            raise GaeError("the optimizer must hold exactly the model's parameters")
        self.dims = [self.Ws[0].shape[1]] + [w.shape[0] for w in self.Ws]
        for p in self.params:                                   # persistent gradient buffers
            p.grad = torch.zeros_like(p)
        self.step_count = 0
        for p in self.params:                                   # torch.optim.Adam's lazy state, created up front
            st = optim.state[p]
            if len(st) == 0:
                st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                st["exp_avg"] = torch.zeros_like(p)
                st["exp_avg_sq"] = torch.zeros_like(p)
            else:
                self.step_count = int(float(st["step"]))
        L = len(self.Ws)
        arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])  # noqa: E731
        self._W, self._b = arr(self.Ws), arr(self.bs)
        self._dW, self._db = arr([w.grad for w in self.Ws]), arr([b.grad for b in self.bs])
        self._p = arr(self.params)
        self._g = arr([p.grad for p in self.params])
        self._m = arr([optim.state[p]["exp_avg"] for p in self.params])
        self._v = arr([optim.state[p]["exp_avg_sq"] for p in self.params])
        self._numel = (ctypes.c_int64 * len(self.params))(*[p.numel() for p in self.params])
        self._desc = StepDesc()
        self._desc.n_layers = L
        for i, v in enumerate(self.dims):
            self._desc.dims[i] = int(v)
        for i, a in enumerate(self.codes):
            self._desc.acts[i] = int(a)
        self._desc.dropout_p = float(model.decoder.dropout)

    def sync_state(self) -> None:
        """Write the step count into torch's state tensors (call before optim.state_dict())."""
        for p in self.params:
            self.optim.state[p]["step"].fill_(float(self.step_count))

    @torch.no_grad()
    def __call__(self, g, per_graph: bool = False, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """loss = model.loss(g); optim.zero_grad(); loss.backward(); optim.step() -- returns the 0-dim loss."""
        lib = _lib.load()
        X = ops.as_rows(g.ndata["h"], "features")
        if X.device != g.device:
            g.to(X.device)
        n, dev = X.shape[0], X.device
        csr, csr_t = g.csr(), g.csr_t()
        desc = self._desc
        desc.pos_weight, desc.per_graph = float(pos_weight_of(g, False, per_graph)), int(bool(per_graph))
        plan = ctypes.byref(csr.plan.struct) if csr.plan is not None and (csr.plan.n_seg or csr.plan.bins) else None
        plan_t = ctypes.byref(csr_t.plan.struct) if csr_t.plan is not None and (csr_t.plan.n_seg or csr_t.plan.bins) else None
        ws_bytes = lib.gae_step_ws_bytes(ctypes.byref(desc), n, plan, plan_t)
        if ws_bytes <= 0:
            raise GaeError("gae_step_ws_bytes rejected the configuration")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        Z = ops.alloc_rows(n, self.dims[-1], dev)
        lo = hi = None
        n_pairs = 0.0
        if per_graph:
            lo, hi, n_pairs = g.block_ranges()
        m = rng = None
        if mask is not None:
            m = mask.to(device=dev, dtype=torch.uint8).contiguous()
        else:
            rng = self.model.decoder._rng_state(dev)
        stream = ops._stream()
        rc = lib.gae_step_fwd_bwd_f32(ctypes.byref(desc), n, ops._ptr(csr.rowptr), ops._ptr(csr.col), plan,
                                      ops._ptr(csr_t.rowptr), ops._ptr(csr_t.col), plan_t, ops._ptr(X), ops._ld(X),
                                      self._W, self._b, ops._ptr(m), ops._ptr(rng), ops._ptr(lo), ops._ptr(hi),
                                      float(n_pairs), 1, ops._ptr(loss), ops._ptr(Z), ops._ld(Z), self._dW, self._db,
                                      ops._ptr(ws), ws_bytes, stream)
        _lib.check(rc, "gae_step_fwd_bwd_f32")
        self.step_count += 1
        b1, b2 = self.group["betas"]
        rc = lib.gae_adam_step_f32(len(self.params), self._p, self._g, self._m, self._v, self._numel,
                                   float(self.group["lr"]), float(b1), float(b2), float(self.group["eps"]),
                                   self.step_count, stream)
        _lib.check(rc, "gae_adam_step_f32")
        g.ndata["h"] = Z                                        # gae.py:53 side effect
        return loss
