"""CPU oracle for the GAE encoder/decoder hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under ``gae_dgl_b200/``
imports it; the product path fails loudly when the CUDA library is missing.

What is restated (all citations relative to ``/root/reference``):

* ``gae_dgl/gae.py:18-19,26-31``  -- ``update_all(copy_src, sum)`` then ``apply_nodes``:
  ``H' = act((A @ H) @ W.T + b)`` with ``A[v, u]`` = number of edges u -> v
  (multigraph, no self loops added, no normalisation).
* ``gae_dgl/gae.py:33-61``        -- layer stack, which layers get ReLU, ``forward``
  (writes the embedding back into ``g.ndata['h']``) and ``encode``.
* ``gae_dgl/gae.py:63-72``        -- decoder: ``F.dropout(z, p)`` with ``training=True``
  ALWAYS (one mask shared by both factors), ``mm(z, z.t())``, identity activation.
* ``gae_dgl/train_inductive.py:44-48`` / ``train_transductive.py:59-65`` -- dense
  adjacency with summed duplicates, ``pos_weight = (N*N - sum(A)) / sum(A)`` and
  ``binary_cross_entropy_with_logits(logits, A, pos_weight=...)`` (mean over N*N).

PARITY STATUS.  Pinned against the reference's own code for everything the reference
contains: ``tests/golden/make_golden_reference.py`` imports ``/root/reference/gae_dgl/gae.py``
unmodified and executes ``class Trainer`` cut out of ``train_inductive.py`` with ``ast``,
and the committed outputs (``tests/golden/ref_gae_steps.npz``: logits, embeddings, adjacency,
pos_weight, per-step loss, gradients, weights after Adam, evaluation loss, four model shapes)
are what ``tests/test_oracle_pins.py`` holds this oracle to.  PARITY UNPINNED for the DGL
primitives only: ``dgl`` is not installed and not installable here (no wheel, no network;
the reference ships no tests or golden vectors), so ``update_all(copy_src, sum)``,
``adjacency_matrix()``, ``dgl.batch`` and ``in_degrees`` run through a small stand-in written
from DGL 0.4's documented semantics in that generator script (degree-bucketed mailbox sum --
deliberately not this module's code).  Every other arithmetic op (``nn.Linear``,
``F.dropout``, ``torch.mm``, BCE-with-logits, Adam) is the very torch function the reference
calls.  The known-answer pins of SURVEY.md section 8c are checked in the same test file.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# Graph indexing (bit-exact integer work)
# --------------------------------------------------------------------------------------


def coo_to_csr(src: torch.Tensor, dst: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """CSR over destination rows: row v lists the sources u of every edge u -> v,
    sorted by (dst, src), duplicates kept (multigraph).  Returns (rowptr int64 [n+1],
    col int32 [E]).  Orientation follows gae.py:18-19 (messages flow src -> dst and are
    summed at dst) and DGL 0.4 ``adjacency_matrix()`` (rows = dst)."""
    src = src.to(torch.int64).cpu()
    dst = dst.to(torch.int64).cpu()
    key = dst * n + src
    order = torch.argsort(key, stable=True)
    col = src[order].to(torch.int32)
    deg = torch.bincount(dst, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(deg, 0)
    return rowptr, col


def csr_transpose(rowptr: torch.Tensor, col: torch.Tensor, n_cols: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """CSR of A^T (rows = sources).  Used by the backward SpMM dX = A^T dY."""
    n = rowptr.numel() - 1
    n_cols = n if n_cols is None else n_cols
    deg = rowptr[1:] - rowptr[:-1]
    dst = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), deg)
    src = col.to(torch.int64)
    # transpose: rows = src, entries = dst
    key = src * n + dst
    order = torch.argsort(key, stable=True)
    colt = dst[order].to(torch.int32)
    degt = torch.bincount(src, minlength=n_cols)
    rowptrt = torch.zeros(n_cols + 1, dtype=torch.int64)
    rowptrt[1:] = torch.cumsum(degt, 0)
    return rowptrt, colt


def in_degrees(rowptr: torch.Tensor) -> torch.Tensor:
    """train_transductive.py:55 -- ``g.in_degrees()`` (int64)."""
    return rowptr[1:] - rowptr[:-1]


def dense_adj(src: torch.Tensor, dst: torch.Tensor, n: int, dtype=torch.float32) -> torch.Tensor:
    """train_inductive.py:44 -- ``g.adjacency_matrix().to_dense()``: A[dst, src] with
    duplicate edges SUMMED (COO -> dense), values 1.0."""
    idx = torch.stack([dst.to(torch.int64), src.to(torch.int64)])
    vals = torch.ones(idx.shape[1], dtype=dtype)
    return torch.sparse_coo_tensor(idx, vals, (n, n)).to_dense()


def dense_adj_from_csr(rowptr: torch.Tensor, col: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    n = rowptr.numel() - 1
    deg = rowptr[1:] - rowptr[:-1]
    dst = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), deg)
    return dense_adj(col.to(torch.int64), dst, n, dtype)


def batch_graphs(graphs: Sequence[Tuple[torch.Tensor, torch.Tensor, int]]):
    """train_inductive.py:34 -- ``dgl.batch``: block-diagonal disjoint union, node ids of
    graph k offset by the prefix sum of the node counts.  graphs = [(src, dst, n), ...]."""
    srcs, dsts, off = [], [], 0
    for s, d, n in graphs:
        srcs.append(s.to(torch.int64) + off)
        dsts.append(d.to(torch.int64) + off)
        off += n
    if not srcs:
        return torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64), 0
    return torch.cat(srcs), torch.cat(dsts), off


# --------------------------------------------------------------------------------------
# Encoder
# --------------------------------------------------------------------------------------


def spmm_sum(rowptr: torch.Tensor, col: torch.Tensor, X: torch.Tensor, vals: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gae.py:18-19,28 -- ``update_all(copy_src('h','m'), sum('m','h'))``:
    Y[v, :] = sum over in-edges (u -> v) of X[u, :]; rows with no in-edge are zero.
    Arithmetic is done in X.dtype (use float64 for ground truth)."""
    n = rowptr.numel() - 1
    deg = rowptr[1:] - rowptr[:-1]
    dst = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), deg)
    msgs = X.index_select(0, col.to(torch.int64))
    if vals is not None:
        msgs = msgs * vals.to(X.dtype)[:, None]
    Y = torch.zeros((n, X.shape[1]), dtype=X.dtype)
    Y.index_add_(0, dst, msgs)
    return Y


def spmm_sum_sparse(rowptr: torch.Tensor, col: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
    """Same result through ``torch.sparse_csr_tensor @ X`` (multi-threaded; this is the
    variant timed as the CPU baseline on the large RMAT workloads)."""
    n = rowptr.numel() - 1
    A = torch.sparse_csr_tensor(rowptr, col.to(torch.int64), torch.ones(col.numel(), dtype=X.dtype),
                                size=(n, X.shape[0]))
    return A @ X


def gcn_layer(rowptr, col, H, W, b, relu: bool):
    """gae.py:26-31 + 13-16 -- aggregate FIRST, then Linear, then activation."""
    Y = spmm_sum(rowptr, col, H)
    out = F.linear(Y, W, b)
    return F.relu(out) if relu else out


def relu_flags(n_layers: int) -> List[bool]:
    """gae.py:36-45 -- every layer but the last gets ReLU; a single layer gets identity."""
    return [i != n_layers - 1 for i in range(n_layers)]


def encode(rowptr, col, X, weights: Sequence[Tuple[torch.Tensor, torch.Tensor]]):
    """gae.py:57-61."""
    h = X
    flags = relu_flags(len(weights))
    for (W, b), r in zip(weights, flags):
        h = gcn_layer(rowptr, col, h, W, b, r)
    return h


# --------------------------------------------------------------------------------------
# Decoder + loss
# --------------------------------------------------------------------------------------


def apply_dropout_mask(z: torch.Tensor, keep_mask: Optional[torch.Tensor], p: float) -> torch.Tensor:
    """gae.py:70 -- ``F.dropout(z, p)`` (training=True by default, so it drops in eval
    mode too).  With an injected keep mask this is exactly what F.dropout computes:
    z * mask / (1 - p)."""
    if keep_mask is None:
        return F.dropout(z, p)
    return z * keep_mask.to(z.dtype) * (1.0 / (1.0 - p))


def decoder_logits(z: torch.Tensor, keep_mask: Optional[torch.Tensor], p: float = 0.1) -> torch.Tensor:
    """gae.py:69-72 with the identity activation GAE installs (gae.py:47)."""
    zd = apply_dropout_mask(z, keep_mask, p)
    return torch.mm(zd, zd.t())


def pos_weight_inductive(adj: torch.Tensor) -> torch.Tensor:
    """train_inductive.py:46 (0-dim tensor, python int minus fp32 tensor)."""
    return (adj.shape[0] * adj.shape[0] - adj.sum()) / adj.sum()


def pos_weight_transductive(adj: torch.Tensor) -> torch.Tensor:
    """train_transductive.py:60 (1-element tensor built from a python float)."""
    return torch.Tensor([float(adj.shape[0] * adj.shape[0] - adj.sum()) / adj.sum()])


def bce_loss(logits: torch.Tensor, adj: torch.Tensor, pos_weight: torch.Tensor) -> torch.Tensor:
    """train_inductive.py:48."""
    return F.binary_cross_entropy_with_logits(logits, adj, pos_weight=pos_weight.to(logits.dtype))


def softplus(x: torch.Tensor) -> torch.Tensor:
    return torch.clamp(x, min=0) + torch.log1p(torch.exp(-x.abs()))


def bce_loss_sparse_form(zd: torch.Tensor, rowptr, col, pos_weight: float) -> torch.Tensor:
    """Closed-form identity, exact for multigraph targets y = multiplicity:
    L = (1/N^2) [ sum_ij softplus(x_ij) + sum_{e=(i,j)} (pw*softplus(-x_ij) - softplus(x_ij)) ]."""
    n = zd.shape[0]
    x = zd @ zd.t()
    dense = softplus(x).sum()
    deg = rowptr[1:] - rowptr[:-1]
    i = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), deg)
    j = col.to(torch.int64)
    xe = (zd[i] * zd[j]).sum(1)
    sparse = (pos_weight * softplus(-xe) - softplus(xe)).sum()
    return (dense + sparse) / float(n * n)


def bce_loss_blockdiag(zd: torch.Tensor, adj: torch.Tensor, sizes: Sequence[int], pos_weight: float) -> torch.Tensor:
    """Per-graph variant (not in the reference): the element-wise weighted
    BCE of train_inductive.py:48 summed over the diagonal blocks only, divided by sum_k n_k^2."""
    total, pairs, lo = zd.new_zeros(()), 0, 0
    pw = torch.as_tensor(pos_weight, dtype=zd.dtype)
    for n_k in sizes:
        hi = lo + n_k
        x = zd[lo:hi] @ zd[lo:hi].t()
        total = total + F.binary_cross_entropy_with_logits(x, adj[lo:hi, lo:hi], pos_weight=pw, reduction="sum")
        pairs += n_k * n_k
        lo = hi
    return total / pairs


def train_step(rowptr, col, X, weights, keep_mask, p=0.1, transductive=False, dtype=torch.float32):
    """One full forward + backward of the reference step (train_inductive.py:44-51).
    Returns (loss, embeddings, [grad W, grad b per layer]).  ``weights`` are (W, b) pairs;
    fresh leaf copies are made in ``dtype``."""
    ws = [(W.detach().to(dtype).clone().requires_grad_(True), b.detach().to(dtype).clone().requires_grad_(True))
          for W, b in weights]
    Xd = X.to(dtype)
    adj = dense_adj_from_csr(rowptr, col, dtype=dtype)
    pw = pos_weight_transductive(adj.float()) if transductive else pos_weight_inductive(adj.float())
    z = encode(rowptr, col, Xd, ws)
    logits = decoder_logits(z, keep_mask, p)
    loss = bce_loss(logits, adj, pw.to(dtype))
    loss.backward()
    grads = [(W.grad.clone(), b.grad.clone()) for W, b in ws]
    return loss.detach(), z.detach(), grads


def _decoder_loss_blocked_backward(zd: torch.Tensor, rowptr, col, pos_weight: float, block: int = 2048):
    """Value of bce_loss_sparse_form and its gradient w.r.t. zd, the dense N x N term evaluated in row
    blocks of `block` (the N = 19 717 Pubmed step does not fit a dense fp64 autograd graph comfortably:
    N^2 = 3.9e8 elements per temporary).  Same arithmetic as bce_loss_sparse_form, block by block."""
    n = zd.shape[0]
    leaf = zd.detach().clone().requires_grad_(True)
    total = 0.0
    for lo in range(0, n, block):
        l_b = softplus(leaf[lo:lo + block] @ leaf.t()).sum() / float(n * n)
        l_b.backward()
        total += float(l_b.detach())
    deg = rowptr[1:] - rowptr[:-1]
    i = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), deg)
    j = col.to(torch.int64)
    xe = (leaf[i] * leaf[j]).sum(1)
    l_s = (pos_weight * softplus(-xe) - softplus(xe)).sum() / float(n * n)
    l_s.backward()
    return total + float(l_s.detach()), leaf.grad


def train_step_blocked(rowptr, col, X, weights, keep_mask, p=0.1, transductive=False, dtype=torch.float64,
                       block: int = 2048):
    """train_step (train_inductive.py:44-51) without any N x N array: the decoder loss in its closed form
    (bce_loss_sparse_form, exact for multigraph targets), dense term in row blocks.  Pinned against
    train_step -- which executes the cited lines literally -- in tests/test_oracle_pins.py."""
    ws = [(W.detach().to(dtype).clone().requires_grad_(True), b.detach().to(dtype).clone().requires_grad_(True))
          for W, b in weights]
    n = rowptr.numel() - 1

    class _Adj:      # what pos_weight_* read of the dense adjacency: its shape and its (fp32, exact < 2^24) sum
        shape = (n, n)

        @staticmethod
        def sum():
            return torch.tensor(float(col.numel()), dtype=torch.float32)

    pw = float(pos_weight_transductive(_Adj) if transductive else pos_weight_inductive(_Adj))
    z = encode(rowptr, col, X.to(dtype), ws)
    zd = apply_dropout_mask(z, keep_mask, p)
    loss, g_zd = _decoder_loss_blocked_backward(zd, rowptr, col, pw, block)
    zd.backward(g_zd)
    grads = [(W.grad.clone(), b.grad.clone()) for W, b in ws]
    return torch.tensor(loss, dtype=dtype), z.detach(), grads


def vgae_train_step(rowptr, col, X, trunk, mu_head, logstd_head, eps, keep_mask, p=0.1, transductive=False,
                    dtype=torch.float64):
    """fp64 restatement of the VGAE step (Kipf & Welling 2016; the reference has no VGAE code, README.md:58
    cites the paper -- the module under test defines the architecture: shared GCN trunk with ReLU, two
    identity GCN heads, z = mu + eps * exp(logstd), the GAE decoder and loss lines (gae.py:69-72,
    train_inductive.py:44-48) plus vgae_kl).  Returns (loss, mu, logstd, grads of [trunk..., mu, logstd])."""
    mk = lambda W, b: (W.detach().to(dtype).clone().requires_grad_(True), b.detach().to(dtype).clone().requires_grad_(True))  # noqa: E731
    ws = [mk(W, b) for W, b in list(trunk) + [mu_head, logstd_head]]
    h = X.to(dtype)
    for W, b in ws[:-2]:
        h = gcn_layer(rowptr, col, h, W, b, relu=True)
    mu = gcn_layer(rowptr, col, h, ws[-2][0], ws[-2][1], relu=False)
    logstd = gcn_layer(rowptr, col, h, ws[-1][0], ws[-1][1], relu=False)
    z = mu + eps.to(dtype) * torch.exp(logstd)
    adj = dense_adj_from_csr(rowptr, col, dtype=dtype)
    pw = pos_weight_transductive(adj.float()) if transductive else pos_weight_inductive(adj.float())
    loss = bce_loss(decoder_logits(z, keep_mask, p), adj, pw.to(dtype)) + vgae_kl(mu, logstd)
    loss.backward()
    return loss.detach(), mu.detach(), logstd.detach(), [(W.grad.clone(), b.grad.clone()) for W, b in ws]


# --------------------------------------------------------------------------------------
# Module-shaped oracle (same constructor / state_dict keys as gae.py) so that parity
# tests read like tests of the reference module.
# --------------------------------------------------------------------------------------


class _NodeApply(nn.Module):  # gae.py:7-16
    def __init__(self, in_feats, out_feats, activation):
        super().__init__()
        self.linear = nn.Linear(in_feats, out_feats)
        self.activation = activation


class _GCN(nn.Module):  # gae.py:21-31
    def __init__(self, in_feats, out_feats, activation):
        super().__init__()
        self.apply_mod = _NodeApply(in_feats, out_feats, activation)

    def forward(self, rowptr, col, feature):
        y = spmm_sum(rowptr, col, feature)
        return self.apply_mod.activation(self.apply_mod.linear(y))


class OracleGAE(nn.Module):
    """gae.py:33-61 on explicit CSR.  ``forward(rowptr, col, X, keep_mask)`` -> [N,N] logits."""

    def __init__(self, in_dim: int, hidden_dims: Sequence[int], dropout: float = 0.1):
        super().__init__()
        ident: Callable = lambda x: x
        dims = [in_dim] + list(hidden_dims)
        flags = relu_flags(len(hidden_dims))
        self.layers = nn.ModuleList(
            [_GCN(dims[i], dims[i + 1], F.relu if flags[i] else ident) for i in range(len(hidden_dims))])
        self.dropout = dropout

    def encode(self, rowptr, col, X):
        h = X
        for conv in self.layers:
            h = conv(rowptr, col, h)
        return h

    def forward(self, rowptr, col, X, keep_mask=None):
        return decoder_logits(self.encode(rowptr, col, X), keep_mask, self.dropout)


def vgae_kl(mu: torch.Tensor, logstd: torch.Tensor) -> torch.Tensor:
    """Kipf & Welling VGAE KL term (no reference code exists -- README.md:58 cites the
    paper only): -0.5/N * mean_i sum_k (1 + 2 logstd - mu^2 - exp(2 logstd))."""
    n = mu.shape[0]
    return -0.5 / n * torch.mean(torch.sum(1 + 2 * logstd - mu.pow(2) - torch.exp(2 * logstd), dim=1))
