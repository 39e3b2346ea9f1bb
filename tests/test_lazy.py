"""CPU: mechanics of the deferred N x N tensors (gae_dgl_b200/lazy.py) with CPU test doubles for the two CUDA
ops behind them.  The literal reference lines (train_inductive.py:44-48) must reach the FUSED loss when
logits and target come from the same graph, and must fall back to real dense tensors on any other use."""
import torch
import torch.nn.functional as F

import gae_dgl_b200 as G
from gae_dgl_b200 import lazy, ops
from oracle import gae_oracle as O


class _Dec:
    dropout = 0.1


def _doubles(monkeypatch, calls):
    def loss_double(z, graph, pw, p, m, st, per_graph):
        calls.append("fused")
        zd = O.apply_dropout_mask(z, m.bool(), p)
        return O.bce_loss(zd @ zd.t(), graph.adjacency_matrix_sparse().to_dense(), torch.tensor(pw))

    def logits_double(z, p, m, st):
        calls.append("dense")
        return O.decoder_logits(z, m.bool(), p)

    monkeypatch.setattr(ops.DecoderLossFunction, "apply", staticmethod(loss_double))
    monkeypatch.setattr(ops.DecoderLogitsFunction, "apply", staticmethod(logits_double))


def _graph():
    g = G.DGLGraph()
    g.add_nodes(6)
    g.add_edges([0, 1, 2, 2, 5], [1, 2, 3, 3, 0])          # one duplicate edge
    return g


def test_lazy_adjacency_answers_from_the_graph():
    g = _graph()
    A = lazy.LazyAdjacency(g)
    assert isinstance(A, torch.Tensor) and A.shape == (6, 6) and A.shape[0] == 6 and A.dim() == 2 and A.size(1) == 6
    assert A.dtype == torch.float32 and A.device.type == "cpu" and A._dense is None
    assert A.to(torch.device("cpu")) is A and A.to("cpu", torch.float32) is A and A.float() is A
    s = A.sum()
    assert type(s) is torch.Tensor and float(s) == 5.0 and A._dense is None           # nothing N x N so far
    pw = (A.shape[0] * A.shape[0] - A.sum()) / A.sum()                                   # train_inductive.py:46
    assert float(pw) == float(O.pos_weight_inductive(g.adjacency_matrix_sparse().to_dense()))
    dense = g.adjacency_matrix_sparse().to_dense()
    assert torch.equal(A + 0, dense) and float(A[3, 2]) == 2.0 and torch.equal(A.sum(0), dense.sum(0))
    assert A.to(torch.float64).dtype == torch.float64                                     # a real conversion materialises
    # the handle: to_dense() deferred, everything else the sparse tensor's
    h = lazy.AdjacencyHandle(g)
    assert isinstance(h.to_dense(), lazy.LazyAdjacency) and h.is_sparse and h._nnz() == 5


def test_literal_reference_lines_route_to_the_fused_loss(monkeypatch):
    calls = []
    _doubles(monkeypatch, calls)
    g = _graph()
    gen = torch.Generator().manual_seed(0)
    z = torch.randn(6, 4, generator=gen, requires_grad=True)
    mask = (torch.rand(6, 4, generator=gen) > 0.1).to(torch.uint8)
    adj = lazy.LazyAdjacency(g).to(torch.device("cpu"))                                   # :44
    pos_weight = (adj.shape[0] * adj.shape[0] - adj.sum()) / adj.sum()                    # :46
    logits = lazy.LazyLogits(z, mask, g, _Dec)                                            # :47
    loss = F.binary_cross_entropy_with_logits(logits, adj, pos_weight=pos_weight)         # :48
    assert calls == ["fused"]
    ref = O.bce_loss(O.decoder_logits(z, mask.bool(), 0.1), g.adjacency_matrix_sparse().to_dense(), pos_weight)
    assert abs(float(loss) - float(ref)) < 1e-6
    loss.backward()
    assert z.grad is not None and float(z.grad.abs().sum()) > 0
    # a different target, a non-default reduction, or a transposed adjacency: dense path, same numbers
    for kw, tgt in (({}, adj + 0), ({"reduction": "sum"}, adj), ({}, lazy.LazyAdjacency(g, transpose=True)),
                    ({}, lazy.LazyAdjacency(_graph()))):
        calls.clear()
        out = F.binary_cross_entropy_with_logits(logits, tgt, pos_weight=pos_weight, **kw)
        assert "fused" not in calls and torch.isfinite(out)
    # any other use of the logits materialises what the eager decoder returns, once
    calls.clear()
    assert torch.equal(torch.sigmoid(logits), torch.sigmoid(O.decoder_logits(z, mask.bool(), 0.1)))
    assert logits.detach().shape == (6, 6) and float(logits[0, 0]) == float(O.decoder_logits(z, mask.bool(), 0.1)[0, 0])
    assert calls.count("dense") <= 1
