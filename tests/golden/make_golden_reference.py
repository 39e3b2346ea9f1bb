"""Generates tests/golden/ref_gae_steps.npz by EXECUTING THE REFERENCE'S OWN CODE.

What runs unmodified, from where it lies under /root/reference (nothing is copied):

* ``gae_dgl/gae.py`` -- imported as a module (NodeApplyModule, GCN, GAE, InnerProductDecoder);
* ``class Trainer`` of ``gae_dgl/train_inductive.py:37-57`` -- the class statement is cut out of
  the file with ``ast`` and executed (the module itself cannot be imported: it parses
  ``sys.argv`` at import time and needs matplotlib / sklearn / dill data files), so the loss
  lines ``:44-48`` and the Adam step ``:50-52`` are the reference's, not a restatement.

What is a stand-in: the ``dgl`` package is not installed and not installable here (no wheel,
no network).  ``_DglShim`` below supplies the few DGL calls those two files make, written from
DGL 0.4's documented semantics and deliberately in a different style from ``oracle/`` so that
the two do not share code:

* ``update_all(copy_src, sum)``: degree bucketing, as DGL 0.4 executes reduce functions --
  nodes are grouped by in-degree k, their mailbox is the ``[nodes, k, d]`` stack of the source
  features along the in-edges, reduced with ``sum(dim=1)``; nodes without in-edges get zeros;
* ``apply_nodes(func)``: ``func(NodeBatch)`` over all nodes, returned fields written to ndata;
* ``adjacency_matrix()``: sparse COO, rows = destination, cols = source, one 1.0 per edge
  (``.to_dense()`` sums duplicates);
* ``dgl.batch``: disjoint union, node ids shifted by the running node count.

So these vectors pin every line of gae.py and of Trainer.iteration against the oracle and the
CUDA path; the DGL primitives themselves stay pinned only by the shim.

Run from the repo root (needs /root/reference; the fixture is committed and travels):
    python tests/golden/make_golden_reference.py
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/gae_dgl"
HERE = os.path.dirname(os.path.abspath(__file__))


# ----------------------------------------------------------------------------------------------
# minimal dgl stand-in
# ----------------------------------------------------------------------------------------------

class _NodeBatch:
    def __init__(self, data):
        self.data = data


class _Graph:
    def __init__(self):
        self._n = 0
        self._src, self._dst = [], []
        self.ndata = {}

    def add_nodes(self, k):
        self._n += int(k)

    def add_edges(self, u, v):
        self._src += [int(x) for x in u]
        self._dst += [int(x) for x in v]

    def number_of_nodes(self):
        return self._n

    def update_all(self, message_func, reduce_func):
        kind_m, src_field, msg_field = message_func
        kind_r, msg_field_r, out_field = reduce_func
        assert kind_m == "copy_src" and kind_r == "sum" and msg_field == msg_field_r
        x = self.ndata[src_field]
        inbox = [[] for _ in range(self._n)]
        for u, v in zip(self._src, self._dst):          # edge-insertion order
            inbox[v].append(u)
        out = torch.zeros((self._n,) + tuple(x.shape[1:]), dtype=x.dtype)
        buckets = {}
        for v, us in enumerate(inbox):
            if us:
                buckets.setdefault(len(us), []).append(v)
        pieces, owners = [], []
        for k, nodes in sorted(buckets.items()):
            idx = torch.tensor([inbox[v] for v in nodes], dtype=torch.int64)       # [nodes, k]
            mailbox = x[idx.reshape(-1)].reshape(len(nodes), k, *x.shape[1:])
            pieces.append(mailbox.sum(dim=1))
            owners += nodes
        if pieces:
            out = out.index_copy(0, torch.tensor(owners, dtype=torch.int64), torch.cat(pieces))
        self.ndata[out_field] = out

    def apply_nodes(self, func):
        self.ndata.update(func(_NodeBatch(self.ndata)))

    def adjacency_matrix(self):
        idx = torch.tensor([self._dst, self._src], dtype=torch.int64).reshape(2, -1)
        return torch.sparse_coo_tensor(idx, torch.ones(len(self._src)), (self._n, self._n))

    def to(self, device):
        return self


def _batch(graphs):
    bg = _Graph()
    feats = []
    for g in graphs:
        off = bg._n
        bg.add_nodes(g._n)
        bg.add_edges([u + off for u in g._src], [v + off for v in g._dst])
        feats.append(g.ndata["h"])
    bg.ndata["h"] = torch.cat(feats)
    return bg


def install_shim():
    dgl = types.ModuleType("dgl")
    fn = types.ModuleType("dgl.function")
    fn.copy_src = lambda src, out: ("copy_src", src, out)
    fn.sum = lambda msg, out: ("sum", msg, out)
    nn_mod = types.ModuleType("dgl.nn")
    nn_pt = types.ModuleType("dgl.nn.pytorch")
    nn_pt.GraphConv = object                      # imported by gae.py:2, never used
    dgl.function, dgl.nn, nn_mod.pytorch = fn, nn_mod, nn_pt
    dgl.DGLGraph, dgl.batch = _Graph, _batch
    sys.modules.update({"dgl": dgl, "dgl.function": fn, "dgl.nn": nn_mod, "dgl.nn.pytorch": nn_pt})
    return dgl


def load_reference():
    install_shim()
    spec = importlib.util.spec_from_file_location("ref_gae", os.path.join(REF, "gae.py"))
    gae = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gae)
    with open(os.path.join(REF, "train_inductive.py")) as f:
        tree = ast.parse(f.read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Trainer"]
    assert len(cls) == 1
    ns = {"torch": torch, "os": os, "device": torch.device("cpu"),
          "BCELoss": F.binary_cross_entropy_with_logits}       # train_inductive.py:10,29
    exec(compile(ast.Module(body=cls, type_ignores=[]), "train_inductive.py::Trainer", "exec"), ns)
    return gae, ns["Trainer"]


# ----------------------------------------------------------------------------------------------
# cases
# ----------------------------------------------------------------------------------------------

def random_multigraph(gen, n, e, hub=None):
    src = torch.randint(0, n, (e,), generator=gen)
    dst = torch.randint(0, n, (e,), generator=gen)
    src = torch.cat([src, src[:5], torch.tensor([2])])       # duplicate edges and a self loop
    dst = torch.cat([dst, dst[:5], torch.tensor([2])])
    if hub is not None:
        k = n // 2
        src = torch.cat([src, torch.randint(0, n, (k,), generator=gen)])
        dst = torch.cat([dst, torch.full((k,), hub)])
    keep = dst != n - 1                                        # last node: no in-edges
    return src[keep], dst[keep]


def mol_like(gen, n):
    """Bidirectional edges as prepare_data.py:61-64 adds them."""
    u = torch.arange(1, n)
    v = (torch.rand(n - 1, generator=gen) * u).long()          # random tree
    return torch.cat([u, v]), torch.cat([v, u])


def build_graph(dgl, src, dst, n, X):
    g = dgl.DGLGraph()
    g.add_nodes(n)
    g.add_edges(src.tolist(), dst.tolist())
    g.ndata["h"] = X
    return g


def dropout_mask(seed, shape):
    """The keep mask F.dropout(z, 0.1) draws under torch.manual_seed(seed) for a z of this shape."""
    torch.manual_seed(seed)
    return F.dropout(torch.ones(shape), 0.1) != 0


def run_case(gae, Trainer, tag, out, graphs, X_list, in_dim, hidden, lr, n_steps, seed):
    dgl = sys.modules["dgl"]
    torch.manual_seed(seed)
    model = gae.GAE(in_dim, hidden)
    out[f"{tag}_hidden"] = np.asarray(hidden)
    out[f"{tag}_lr"] = np.float64(lr)
    for k, v in model.state_dict().items():
        out[f"{tag}_init.{k}"] = v.numpy().copy()
    trainer = Trainer(model, types.SimpleNamespace(lr=lr))
    members = [build_graph(dgl, s, d, n, X) for (s, d, n), X in zip(graphs, X_list)]
    out[f"{tag}_sizes"] = np.asarray([n for _, _, n in graphs])
    for i, ((s, d, n), X) in enumerate(zip(graphs, X_list)):
        out[f"{tag}_src{i}"], out[f"{tag}_dst{i}"], out[f"{tag}_X{i}"] = s.numpy(), d.numpy(), X.numpy()
    losses = []
    for step in range(n_steps):
        for g, X in zip(members, X_list):
            g.ndata["h"] = X                      # gae.py:53 overwrote it with the embedding
        bg = dgl.batch(members) if len(members) > 1 else members[0]
        n_total = bg.number_of_nodes()
        step_seed = 1000 * seed + step
        mask = dropout_mask(step_seed, (n_total, hidden[-1]))
        out[f"{tag}_mask{step}"] = mask.numpy()
        if step == 0:
            # forward alone (gae.py:49-55): logits, and the embedding written back into ndata['h']
            torch.manual_seed(step_seed)
            with torch.no_grad():
                logits = model.forward(bg)
            emb = bg.ndata["h"]
            zd = emb * mask / 0.9
            assert torch.allclose(zd @ zd.t(), logits, atol=1e-6), "recovered dropout mask is wrong"
            out[f"{tag}_logits"], out[f"{tag}_emb"] = logits.numpy(), emb.numpy()
            with torch.no_grad():
                for g, X in zip(members, X_list):
                    g.ndata["h"] = X
                bg = dgl.batch(members) if len(members) > 1 else members[0]
                out[f"{tag}_encode"] = model.encode(bg).numpy()
            adj = bg.adjacency_matrix().to_dense()
            out[f"{tag}_adj"] = adj.numpy()
            out[f"{tag}_pos_weight"] = ((adj.shape[0] * adj.shape[0] - adj.sum()) / adj.sum()).numpy()
            out[f"{tag}_in_deg"] = adj.sum(1).long().numpy()
            # encode() leaves the graph WITHOUT ndata['h'] (GCN.forward pops it, gae.py:30, and only
            # forward() writes it back, gae.py:53): restore the input features before the step
            assert "h" not in bg.ndata
            for g, X in zip(members, X_list):
                g.ndata["h"] = X
            bg = dgl.batch(members) if len(members) > 1 else members[0]
        torch.manual_seed(step_seed)
        losses.append(trainer.iteration(bg, train=True))          # train_inductive.py:43-53
        for k, v in model.state_dict().items():
            out[f"{tag}_after{step}.{k}"] = v.numpy().copy()
        for k, v in model.named_parameters():                     # left in place by :50-52
            out[f"{tag}_grad{step}.{k}"] = v.grad.numpy().copy()
    out[f"{tag}_losses"] = np.asarray(losses, dtype=np.float64)
    # evaluation call: no step, dropout still active (train_inductive.py:100-105)
    for g, X in zip(members, X_list):
        g.ndata["h"] = X
    bg = dgl.batch(members) if len(members) > 1 else members[0]
    ev_seed = 1000 * seed + 999
    out[f"{tag}_mask_eval"] = dropout_mask(ev_seed, (bg.number_of_nodes(), hidden[-1])).numpy()
    torch.manual_seed(ev_seed)
    model.eval()
    out[f"{tag}_loss_eval"] = np.float64(trainer.iteration(bg, train=False))
    print(tag, "losses", losses, "eval", float(out[f"{tag}_loss_eval"]))


def main():
    gae, Trainer = load_reference()
    gen = torch.Generator().manual_seed(2024)
    out = {}
    # A: one directed multigraph with duplicates, a self loop, a hub row and an isolated node; ZINC dims
    n = 83
    s, d = random_multigraph(gen, n, 320, hub=11)
    run_case(gae, Trainer, "A", out, [(s, d, n)], [torch.randn(n, 39, generator=gen)], 39, [32, 16], 1e-3, 3, seed=1)
    # B: a dgl.batch of five molecule-like graphs (bidirectional edges), one-hot-ish features
    graphs, feats = [], []
    for n_k in (7, 12, 9, 15, 6):
        s, d = mol_like(gen, n_k)
        graphs.append((s, d, n_k))
        feats.append((torch.rand(n_k, 39, generator=gen) < 0.12).float())
    run_case(gae, Trainer, "B", out, graphs, feats, 39, [32, 16], 1e-3, 2, seed=2)
    # C: single layer -> identity activation only (gae.py:44-45); D: three layers (ReLU, ReLU, identity)
    n = 40
    s, d = random_multigraph(gen, n, 150)
    Xc = torch.randn(n, 20, generator=gen)
    run_case(gae, Trainer, "C", out, [(s, d, n)], [Xc], 20, [16], 1e-2, 2, seed=3)
    run_case(gae, Trainer, "D", out, [(s, d, n)], [Xc], 20, [24, 12, 8], 1e-2, 2, seed=4)
    np.savez_compressed(os.path.join(HERE, "ref_gae_steps.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_gae_steps.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
