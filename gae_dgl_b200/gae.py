"""Drop-in for /root/reference/gae_dgl/gae.py on the B200 kernels.

Same classes, constructor signatures, attribute names and state_dict keys
(`layers.{i}.apply_mod.linear.{weight,bias}`) as the reference, so checkpoints written by
either load into the other.  Differences, all additive:
  * `GAE.loss(g)` / `GAE.reconstruction_loss(g)`: the fused decoder + weighted BCE path that
    never materialises N x N (what the trainers in this package call);
  * `InnerProductDecoder.forward(z, mask=None)` accepts an injected keep-mask (parity tests);
  * `VGAE` (not present in the reference, which only cites the paper, README.md:58).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lazy, ops
from .graph import DGLGraph, function as fn

_IDENTITY = lambda x: x  # noqa: E731  (gae.py:43,45,47 use `lambda x:x`)


def _act_code(activation) -> Optional[int]:
    if activation in (F.relu, torch.relu):
        return ops.ACT_RELU
    if activation is _IDENTITY or activation is None:
        return ops.ACT_IDENTITY
    return None  # arbitrary callable: applied after an identity-fused linear


class NodeApplyModule(nn.Module):
    """gae.py:7-16 -- Linear + activation applied to the node frame."""

    def __init__(self, in_feats, out_feats, activation):
        super().__init__()
        self.linear = nn.Linear(in_feats, out_feats)
        self.activation = activation

    def forward(self, node):
        code = _act_code(self.activation)
        h = ops.LinearActFunction.apply(node.data['h'], self.linear.weight, self.linear.bias,
                                        ops.ACT_IDENTITY if code is None else code)
        if code is None:
            h = self.activation(h)
        return {'h': h}


gcn_msg = fn.copy_src(src='h', out='m')      # gae.py:18
gcn_reduce = fn.sum(msg='m', out='h')        # gae.py:19  (sum aggregation)


class GCN(nn.Module):
    """gae.py:21-31 -- aggregate first (SpMM), then NodeApplyModule."""

    def __init__(self, in_feats, out_feats, activation):
        super().__init__()
        self.apply_mod = NodeApplyModule(in_feats, out_feats, activation)

    def forward(self, g, feature):
        # the node frame carries the data through both calls and is emptied of 'h' again on the way
        # out (gae.py:30 pops it: after encode() the graph has no 'h' until the caller sets one)
        frame = g.ndata
        frame['h'] = feature
        g.update_all(gcn_msg, gcn_reduce)          # K1: Y = A H
        g.apply_nodes(func=self.apply_mod)         # K3: act(Y W^T + b)
        return frame.pop('h')


def _build_layers(in_dim: int, hidden_dims: Sequence[int]):
    """gae.py:35-45: ReLU on every layer but the last, identity on the last (a single layer is the last).

    The reference builds `GCN(in_dim, hidden_dims[0], F.relu)` once more than it keeps (:35 is
    overwritten by :37 or :45), which draws one extra set of first-layer initial weights from torch's
    generator.  The same draw is made here, so `torch.manual_seed(s); GAE(...)` starts from bit-identical
    weights in both code bases (pinned by tests/test_oracle_pins.py against the reference run)."""
    nn.Linear(in_dim, hidden_dims[0])              # the discarded draw
    dims = [in_dim] + list(hidden_dims)
    last = len(hidden_dims) - 1
    return [GCN(dims[i], dims[i + 1], _IDENTITY if i == last else F.relu) for i in range(len(hidden_dims))]


class InnerProductDecoder(nn.Module):
    """gae.py:63-72.  Dropout is applied with training=True ALWAYS (F.dropout default), one
    mask shared by both factors of z z^T."""

    def __init__(self, activation=torch.sigmoid, dropout=0.1):
        super().__init__()
        self.dropout = dropout
        self.activation = activation

    def _rng_state(self, device) -> torch.Tensor:
        """Device-resident Philox state {seed, offset}, created on first use from torch's host
        generator (so torch.manual_seed controls it) and advanced on-stream by every draw: no
        per-step host RNG call, and captured CUDA graphs draw a fresh mask on each replay."""
        st = getattr(self, "_philox", None)
        if st is None:
            seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
            st = torch.tensor([int(seed), 0], dtype=torch.int64, device=device)
            self._philox = st
        elif st.device != device:
            st = self._philox = st.to(device)          # the stream continues where it was (no re-seeding)
        return st

    def rng_state_dict(self) -> Optional[torch.Tensor]:
        """{seed, offset} of the dropout stream (host copy) for checkpoints; None before the first draw."""
        st = getattr(self, "_philox", None)
        return None if st is None else st.detach().cpu().clone()

    def load_rng_state(self, state: Optional[torch.Tensor], device=None) -> None:
        if state is not None:
            self._philox = state.to(device if device is not None else state.device).clone()

    def forward(self, z, mask: Optional[torch.Tensor] = None):
        st = None if mask is not None else self._rng_state(z.device)
        x = ops.DecoderLogitsFunction.apply(z, float(self.dropout), mask, st)
        return self.activation(x)

    def loss(self, z, g: DGLGraph, pos_weight: float, mask: Optional[torch.Tensor] = None, per_graph: bool = False):
        """Fused path: mean BCE-with-logits(z_d z_d^T, A, pos_weight) without the N x N arrays
        (replaces gae.py:71 + train_inductive.py:44,48).  per_graph=True restricts the pairs to
        each member graph of a batch (block-diagonal; an extension, not the reference default)."""
        if z.shape[1] > ops.MAX_FUSED_DECODER_WIDTH:
            # wider than the fused kernels go (the reference takes any --hidden_dims; its HPO script searches up
            # to 256): the reference's own formulation -- materialised logits, dense adjacency, torch's BCE
            if per_graph:
                raise ops.GaeError(f"the per-graph decoder needs an embedding width <= {ops.MAX_FUSED_DECODER_WIDTH}")
            adj = g.adjacency_matrix().to_dense().to(z.device)               # train_inductive.py:44
            pw = torch.tensor(float(pos_weight), dtype=z.dtype, device=z.device)
            st = None if mask is not None else self._rng_state(z.device)
            logits = ops.DecoderLogitsFunction.apply(z, float(self.dropout), mask, st)
            return F.binary_cross_entropy_with_logits(logits, adj, pos_weight=pw)           # :47-48
        st = None if mask is not None else self._rng_state(z.device)
        return ops.DecoderLossFunction.apply(z, g, float(pos_weight), float(self.dropout), mask, st, per_graph)


def pos_weight_of(g: DGLGraph, transductive: bool = False, per_graph: bool = False) -> float:
    """train_inductive.py:46 / train_transductive.py:60 evaluated from integers: the sum of the
    dense adjacency equals the edge count (with multiplicity), exactly representable in fp32
    below 2^24 edges; the arithmetic below repeats the reference's fp32 tensor ops.
    per_graph: the pair count is sum_k n_k^2 (block-diagonal decoder) instead of N^2."""
    n = g.number_of_nodes()
    # numpy float32 scalars perform the same IEEE single-precision operations as the 0-dim fp32
    # tensors of the reference expression (checked against torch in tests/test_host_logic.py) at a
    # fraction of the host cost -- this runs once per step in the inductive loop
    adj_sum = np.float32(g.number_of_edges())
    if per_graph:
        pairs = g.block_ranges()[2]
        return float((np.float32(pairs) - adj_sum) / adj_sum)
    if transductive:
        # `python_float / tensor` is Tensor.__rtruediv__ = tensor.reciprocal() * python_float: an fp32
        # reciprocal and an fp32 multiply (train_transductive.py:60), not an fp32 division
        return float((np.float32(1.0) / adj_sum) * (np.float32(n * n) - adj_sum))
    return float((np.float32(n * n) - adj_sum) / adj_sum)


class GAE(nn.Module):
    """gae.py:33-61."""

    def __init__(self, in_dim, hidden_dims):
        super().__init__()
        self.layers = nn.ModuleList(_build_layers(in_dim, hidden_dims))
        self.decoder = InnerProductDecoder(activation=_IDENTITY)

    def encode(self, g):
        """gae.py:57-61: embeddings; the graph is left without ndata['h'] (GCN.forward pops it)."""
        z = g.ndata['h']
        for layer in self.layers:
            z = layer(g, z)
        return z

    def forward(self, g):
        """gae.py:49-55: logits [N, N]; side effect kept: the input features in ndata['h'] are replaced
        by the embeddings (:53)."""
        z = self.encode(g)
        g.ndata['h'] = z
        if lazy.ENABLED and z.is_cuda and self.decoder.activation is _IDENTITY:
            # deferred logits: BCELoss(adj_logits, adj, pos_weight) on them is the fused decoder kernel; any
            # other use materialises exactly what self.decoder(z) returns (same keep-mask)
            return lazy.LazyLogits(z, lazy.keep_mask_for(self.decoder, z), g, self.decoder)
        return self.decoder(z)

    def loss(self, g, pos_weight: Optional[float] = None, mask: Optional[torch.Tensor] = None,
             transductive: bool = False, per_graph: bool = False, fused_step: Optional[bool] = None,
             aggregated_input: bool = False):
        """Fused equivalent of `BCELoss(model.forward(g), adj, pos_weight)`
        (train_inductive.py:44-48), including the gae.py:53 write-back.

        fused_step (default: automatic) runs the whole step -- encoder, decoder loss and the
        backward -- behind ONE native call (ops.FusedStepFunction); it applies when every layer
        activation is ReLU / identity and the input features need no gradient.  The layer-by-layer
        path (what `forward` / `encode` use) computes the same numbers.

        aggregated_input=True: ndata['h'] already holds A X (ops.spmm of the input features, computed once by
        the caller because features and graph do not change between epochs, train_transductive.py:45-46,63);
        the first layer then skips its aggregation -- same bits, one 500-wide SpMM less per epoch."""
        if pos_weight is None:
            pos_weight = pos_weight_of(g, transductive, per_graph)
        x = g.ndata['h']
        codes = [_act_code(conv.apply_mod.activation) for conv in self.layers]
        can_fuse = all(c is not None for c in codes) and not x.requires_grad and \
            self.layers[-1].apply_mod.linear.out_features <= ops.MAX_FUSED_DECODER_WIDTH and len(self.layers) <= 8
        if fused_step is None:
            fused_step = can_fuse
        if aggregated_input and not fused_step:
            raise ops.GaeError("aggregated_input needs the fused step (ReLU/identity activations, d_last <= 64)")
        if fused_step:
            if not can_fuse:
                raise ops.GaeError("fused_step=True needs ReLU/identity activations, leaf features and d_last <= 64")
            if x.device != g.device:
                g.to(x.device)
            dims = [self.layers[0].apply_mod.linear.in_features] + [c.apply_mod.linear.out_features for c in self.layers]
            params = []
            for conv in self.layers:
                params += [conv.apply_mod.linear.weight, conv.apply_mod.linear.bias]
            st = None if mask is not None else self.decoder._rng_state(x.device)
            need_grad = torch.is_grad_enabled() and any(t.requires_grad for t in params)
            loss, z = ops.FusedStepFunction.apply(x, g, dims, codes, float(pos_weight), float(self.decoder.dropout), mask,
                                                  st, (per_graph, aggregated_input), need_grad, *params)
            g.ndata['h'] = z
            return loss
        h = self.encode(g)
        g.ndata['h'] = h
        return self.decoder.loss(h, g, pos_weight, mask, per_graph)

    reconstruction_loss = loss


class VGAE(nn.Module):
    """Variational GAE (Kipf & Welling 2016), same constructor as GAE.  Not in the reference
    (README.md:58 cites the paper only): shared GCN trunk over hidden_dims[:-1], two GCN heads
    (mu, log sigma) of width hidden_dims[-1], reparameterisation, inner-product decoder.
    `loss(g)` returns reconstruction + KL."""

    def __init__(self, in_dim, hidden_dims):
        super().__init__()
        hidden_dims = list(hidden_dims)
        if len(hidden_dims) >= 2:
            trunk = _build_layers(in_dim, hidden_dims[:-1] + [hidden_dims[-1]])[:-1]
            head_in = hidden_dims[-2]
        else:
            trunk, head_in = [], in_dim
        self.layers = nn.ModuleList(trunk)
        self.mu_head = GCN(head_in, hidden_dims[-1], _IDENTITY)
        self.logstd_head = GCN(head_in, hidden_dims[-1], _IDENTITY)
        self.decoder = InnerProductDecoder(activation=_IDENTITY)

    def encode_dist(self, g):
        h = g.ndata['h']
        for conv in self.layers:
            h = conv(g, h)
        return self.mu_head(g, h), self.logstd_head(g, h)

    def encode(self, g):
        return self.encode_dist(g)[0]

    def _sample(self, mu, logstd, eps=None):
        if not self.training and eps is None:
            return mu
        if eps is None:
            eps = torch.randn_like(mu)
        return mu + eps * torch.exp(logstd)

    @staticmethod
    def kl(mu, logstd):
        n = mu.shape[0]
        return -0.5 / n * torch.mean(torch.sum(1 + 2 * logstd - mu.pow(2) - torch.exp(2 * logstd), dim=1))

    def forward(self, g, eps=None):
        mu, logstd = self.encode_dist(g)
        z = self._sample(mu, logstd, eps)
        g.ndata['h'] = z
        return self.decoder(z)

    def loss(self, g, pos_weight: Optional[float] = None, mask=None, eps=None, transductive=False):
        mu, logstd = self.encode_dist(g)
        z = self._sample(mu, logstd, eps)
        g.ndata['h'] = z
        if pos_weight is None:
            pos_weight = pos_weight_of(g, transductive)
        return self.decoder.loss(z, g, pos_weight, mask) + self.kl(mu, logstd)
