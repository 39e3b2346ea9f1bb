"""Deferred N x N tensors, so that the reference's LITERAL training lines land on the fused kernels.

train_inductive.py:44-48 (and train_transductive.py:59-65) read

    adj        = g.adjacency_matrix().to_dense().to(device)
    pos_weight = (adj.shape[0] * adj.shape[0] - adj.sum()) / adj.sum()
    adj_logits = model.forward(g)
    loss       = BCELoss(adj_logits, adj, pos_weight=pos_weight)

Executed as written that is three N x N fp32 arrays per step.  Here `to_dense()` and `forward(g)` return
tensor subclasses WITHOUT storage that remember where they came from:

  * `LazyAdjacency`  -- knows its graph; `.shape`, `.to(same device)`, `.sum()` (= the edge count, exactly
    what the dense sum gives below 2^24 edges) are answered from the graph;
  * `LazyLogits`     -- holds the embeddings and the dropout keep-mask already drawn for this forward pass.

`F.binary_cross_entropy_with_logits(LazyLogits, LazyAdjacency of the same graph, pos_weight=...)` with
the default mean reduction is routed to the fused decoder (ops.DecoderLossFunction: loss and gradient in one
pass, nothing N x N).  ANY other use -- indexing, arithmetic, printing, another loss, a different target --
materialises the real dense tensor first (same numbers the eager path produces: the logits through
ops.DecoderLogitsFunction, so autograd still reaches the embeddings) and re-dispatches the call on it.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import ops

ENABLED = True        # set False to get the eager N x N tensors back everywhere

_META = {"shape", "dtype", "device", "requires_grad", "is_cuda", "ndim", "is_sparse", "layout", "grad_fn", "is_leaf",
         "names", "is_quantized", "is_meta", "_version", "grad", "is_nested", "is_complex", "is_floating_point",
         "size", "dim", "numel", "stride", "is_contiguous", "storage_offset", "element_size", "nelement", "ndimension",
         "__len__", "__hash__", "__class__", "__dir__", "__reduce_ex__", "type"}


def _name_of(func) -> str:
    return getattr(func, "__name__", None) or getattr(getattr(func, "__self__", None), "__name__", "") or str(func)


class _Deferred(torch.Tensor):
    """Common part: a wrapper subclass (metadata only) that materialises on first real use."""

    @staticmethod
    def _wrap(cls, n: int, dtype, device):
        return torch.Tensor._make_wrapper_subclass(cls, (n, n), dtype=dtype, device=device, requires_grad=False)

    def materialize(self) -> torch.Tensor:
        raise NotImplementedError

    def _intercept(self, func, name, args, kwargs):
        return NotImplemented

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = _name_of(func)
        me = next((a for a in args if isinstance(a, _Deferred)), None)
        if name in ("__get__", "__repr__") and me is not None and len(args) == 1 and not kwargs:
            # attribute getters (Tensor.shape.__get__ ...): answer metadata from the wrapper itself
            owner = getattr(func, "__self__", None)
            attr = getattr(owner, "__name__", "")
            if attr in _META:
                with torch._C.DisableTorchFunctionSubclass():
                    return func(*args, **kwargs)
        if name in _META and me is not None:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        for a in args:
            if isinstance(a, _Deferred):
                out = a._intercept(func, name, args, kwargs)
                if out is not NotImplemented:
                    return out
        # anything else: real tensors, then the ordinary call
        conv = lambda x: x.materialize() if isinstance(x, _Deferred) else x  # noqa: E731
        args = tuple(conv(a) for a in args)
        kwargs = {k: conv(v) for k, v in kwargs.items()}
        with torch._C.DisableTorchFunctionSubclass():
            return func(*args, **kwargs)


    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        """Last resort (an ATen call that bypassed __torch_function__): real tensors, ordinary call."""
        import torch.utils._pytree as pytree
        conv = lambda x: x.materialize() if isinstance(x, _Deferred) else x  # noqa: E731
        args, kwargs = pytree.tree_map(conv, (args, kwargs or {}))
        return func(*args, **kwargs)


class LazyAdjacency(_Deferred):
    """`g.adjacency_matrix().to_dense()` without the N x N array."""

    @staticmethod
    def __new__(cls, graph, transpose: bool = False):
        r = _Deferred._wrap(cls, graph.number_of_nodes(), torch.float32, graph.csr().rowptr.device)
        r._graph, r._transpose, r._dense = graph, bool(transpose), None
        return r

    def materialize(self) -> torch.Tensor:
        if self._dense is None:
            self._dense = self._graph.adjacency_matrix_sparse(self._transpose).to_dense()
        return self._dense

    def _intercept(self, func, name, args, kwargs):
        if args and args[0] is self:
            if name in ("to", "cuda", "float", "contiguous", "detach", "clone") and self._stays(name, args[1:], kwargs):
                return self
            if name == "sum" and len(args) == 1 and not kwargs and self._graph.number_of_edges() < (1 << 24):
                # fp32 sum of E ones and N^2 - E zeros: exact below 2^24
                return torch.tensor(float(self._graph.number_of_edges()), dtype=torch.float32, device=self.device)
        return NotImplemented

    def _stays(self, name, rest, kwargs) -> bool:
        if name in ("float", "contiguous", "detach"):
            return not rest and not kwargs
        if name == "clone":
            return not rest and not kwargs
        if name == "cuda":
            dev = rest[0] if rest else kwargs.get("device", None)
            return dev is None or torch.device("cuda", dev if isinstance(dev, int) else torch.device(dev).index or 0) == self.device
        # .to(device) / .to(dtype) / .to(device, dtype)
        for v in list(rest) + [kwargs.get("device"), kwargs.get("dtype")]:
            if v is None or isinstance(v, bool):
                continue
            if isinstance(v, torch.dtype):
                if v != torch.float32:
                    return False
            else:
                try:
                    d = torch.device(v)
                except (TypeError, RuntimeError):
                    return False
                if d.type != self.device.type or (d.index is not None and d.index != self.device.index):
                    return False
        return True


class LazyLogits(_Deferred):
    """`model.forward(g)` (gae.py:49-55) without the N x N array: embeddings + the keep-mask of this pass."""

    @staticmethod
    def __new__(cls, z: torch.Tensor, keep_mask: torch.Tensor, graph, decoder):
        r = _Deferred._wrap(cls, z.shape[0], z.dtype, z.device)
        r._z, r._mask, r._graph, r._decoder, r._dense = z, keep_mask, graph, decoder, None
        return r

    def materialize(self) -> torch.Tensor:
        if self._dense is None:
            self._dense = ops.DecoderLogitsFunction.apply(self._z, float(self._decoder.dropout), self._mask, None)
        return self._dense

    def _intercept(self, func, name, args, kwargs):
        if func is F.binary_cross_entropy_with_logits or name == "binary_cross_entropy_with_logits":
            names = ("input", "target", "weight", "size_average", "reduce", "reduction", "pos_weight")
            b = dict(zip(names, args))
            b.update(kwargs)
            tgt, pw = b.get("target"), b.get("pos_weight")
            plain = b.get("weight") is None and b.get("size_average") is None and b.get("reduce") is None and \
                b.get("reduction", "mean") == "mean"
            if b.get("input") is self and plain and isinstance(tgt, LazyAdjacency) and tgt._graph is self._graph and \
                    not tgt._transpose and self._z.shape[1] <= ops.MAX_FUSED_DECODER_WIDTH and \
                    (pw is None or (isinstance(pw, torch.Tensor) and pw.numel() == 1) or isinstance(pw, (int, float))):
                w = 1.0 if pw is None else float(pw)
                return ops.DecoderLossFunction.apply(self._z, self._graph, w, float(self._decoder.dropout), self._mask,
                                                     None, False)
        return NotImplemented


class AdjacencyHandle:
    """What `DGLGraph.adjacency_matrix()` returns: `.to_dense()` is deferred (LazyAdjacency); every other
    attribute is the real torch sparse COO tensor's (built on first use)."""

    def __init__(self, graph, transpose: bool = False):
        self._graph, self._transpose, self._sparse = graph, transpose, None

    def _real(self) -> torch.Tensor:
        if self._sparse is None:
            self._sparse = self._graph.adjacency_matrix_sparse(self._transpose)
        return self._sparse

    def to_dense(self) -> torch.Tensor:
        return LazyAdjacency(self._graph, self._transpose)

    def __getattr__(self, item):
        return getattr(self._real(), item)

    def __repr__(self):
        return repr(self._real())


def keep_mask_for(decoder, z: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The keep-mask of one decoder pass: the injected one, or a fresh draw from the decoder's device RNG."""
    if mask is not None:
        return mask.to(device=z.device, dtype=torch.uint8).contiguous()
    _, m = ops.dropout_fwd(z.detach(), float(decoder.dropout), None, rng_state=decoder._rng_state(z.device))
    return m
