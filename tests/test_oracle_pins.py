"""CPU: pins the oracle (SURVEY.md section 8c known-answer tests + committed golden vectors)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import gae_oracle as O
from oracle import c_spmm

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "gae_small.npz")


def test_pin_dropout_always_on_and_scaled():
    # pin (1): F.dropout(z, 0.1) with default args drops ~10 % and scales by 1/0.9, train or eval
    torch.manual_seed(0)
    z = torch.ones(200, 50)
    out = F.dropout(z, 0.1)
    kept = out != 0
    assert 0.85 < kept.float().mean() < 0.95
    assert torch.allclose(out[kept], torch.full_like(out[kept], 1 / 0.9))
    mask = kept
    assert torch.equal(O.apply_dropout_mask(z, mask, 0.1), out)


def test_pin_coo_to_dense_sums_duplicates():
    # pin (2): edges {(0,1),(0,1),(1,0)} as (dst,src) index pairs -> [[0,2],[1,0]]
    src = torch.tensor([1, 1, 0])
    dst = torch.tensor([0, 0, 1])
    assert O.dense_adj(src, dst, 2).tolist() == [[0.0, 2.0], [1.0, 0.0]]


def test_pin_sparse_form_identity_and_gradient():
    # pin (3): closed form == F.binary_cross_entropy_with_logits(..., pos_weight), incl. multigraph y=2
    g = torch.Generator().manual_seed(3)
    n = 40
    src = torch.randint(0, n, (150,), generator=g)
    dst = torch.randint(0, n, (150,), generator=g)
    src = torch.cat([src, src[:10]])
    dst = torch.cat([dst, dst[:10]])            # duplicates -> y = 2
    rowptr, col = O.coo_to_csr(src, dst, n)
    adj = O.dense_adj(src, dst, n, torch.float64)
    assert adj.max() >= 2
    z = torch.randn(n, 16, generator=g, dtype=torch.float64, requires_grad=True)
    pw = float(O.pos_weight_inductive(adj))
    ref = F.binary_cross_entropy_with_logits(z @ z.t(), adj, pos_weight=torch.tensor(pw, dtype=torch.float64))
    alt = O.bce_loss_sparse_form(z, rowptr, col, pw)
    assert abs(float(ref) - float(alt)) < 1e-12 * max(1.0, abs(float(ref)))
    g1, = torch.autograd.grad(ref, z, retain_graph=True)
    g2, = torch.autograd.grad(alt, z)
    assert torch.allclose(g1, g2, rtol=1e-10, atol=1e-14)
    # analytic gradient of SURVEY.md 8a row 6
    x = (z @ z.t()).detach()
    s = torch.sigmoid(x)
    G = ((1 - adj) * s - pw * adj * (1 - s)) / (n * n)
    assert torch.allclose(g1, (G + G.t()) @ z.detach(), rtol=1e-9, atol=1e-13)


def test_pin_blocked_train_step_equals_the_literal_one():
    """train_step_blocked (no N x N array; what the Pubmed-size GPU parity test uses) against train_step, which
    executes train_inductive.py:44-51 literally -- several row blocks, duplicate edges, both pos_weight forms."""
    g = torch.Generator().manual_seed(5)
    n, e, f = 157, 900, 23
    src, dst = torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g)
    src, dst = torch.cat([src, src[:40]]), torch.cat([dst, dst[:40]])
    rowptr, col = O.coo_to_csr(src, dst, n)
    X = torch.randn(n, f, generator=g)
    torch.manual_seed(3)
    ref = O.OracleGAE(f, [12, 8])
    weights = [(l.apply_mod.linear.weight.detach(), l.apply_mod.linear.bias.detach()) for l in ref.layers]
    mask = torch.rand(n, 8, generator=g) >= 0.1
    for transductive in (False, True):
        l0, z0, g0 = O.train_step(rowptr, col, X, weights, mask, transductive=transductive, dtype=torch.float64)
        l1, z1, g1 = O.train_step_blocked(rowptr, col, X, weights, mask, transductive=transductive, block=50)
        assert abs(float(l0) - float(l1)) < 1e-6 * abs(float(l0))        # pos_weight is an fp32 value in both
        assert torch.equal(z0, z1)
        for (a, b), (c, d) in zip(g0, g1):
            assert float((a - c).abs().max()) < 1e-9 * max(float(a.abs().max()), 1.0)
            assert float((b - d).abs().max()) < 1e-9 * max(float(b.abs().max()), 1.0)


def test_pin_linear_init_and_param_counts():
    # pins (4), (5)
    lin = nn.Linear(39, 32)
    assert lin.weight.abs().max() <= 1 / np.sqrt(39) + 1e-7
    count = lambda m: sum(p.numel() for p in m.parameters())
    assert count(O.OracleGAE(39, [32, 16])) == 1808
    assert count(O.OracleGAE(500, [32, 16])) == 16560
    assert count(O.OracleGAE(1433, [32, 16])) == 46416
    assert O.relu_flags(1) == [False] and O.relu_flags(3) == [True, True, False]
    assert list(O.OracleGAE(39, [32, 16]).state_dict().keys()) == [
        "layers.0.apply_mod.linear.weight", "layers.0.apply_mod.linear.bias",
        "layers.1.apply_mod.linear.weight", "layers.1.apply_mod.linear.bias"]


def test_spmm_hand_graph():
    # edges u->v: 0->1, 0->1 (dup), 2->1, 1->0 ; node 2 has no in-edge
    src = torch.tensor([0, 0, 2, 1])
    dst = torch.tensor([1, 1, 1, 0])
    rowptr, col = O.coo_to_csr(src, dst, 3)
    assert rowptr.tolist() == [0, 1, 4, 4] and col.tolist() == [1, 0, 0, 2]
    X = torch.tensor([[1.0, 10.0], [2.0, 20.0], [4.0, 40.0]])
    Y = O.spmm_sum(rowptr, col, X)
    assert Y.tolist() == [[2.0, 20.0], [6.0, 60.0], [0.0, 0.0]]
    assert torch.equal(O.dense_adj(src, dst, 3) @ X, Y)
    assert O.in_degrees(rowptr).tolist() == [1, 3, 0]
    rt, ct = O.csr_transpose(rowptr, col)
    assert rt.tolist() == [0, 2, 3, 4] and ct.tolist() == [1, 1, 0, 1]


def test_spmm_variants_agree():
    g = torch.Generator().manual_seed(5)
    n, e = 300, 3000
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n, (e,), generator=g)
    rowptr, col = O.coo_to_csr(src, dst, n)
    X = torch.randn(n, 24, generator=g)
    y64 = O.spmm_sum(rowptr, col, X.double())
    assert torch.allclose(O.spmm_sum(rowptr, col, X).double(), y64, atol=1e-4)
    assert torch.allclose(O.spmm_sum_sparse(rowptr, col, X).double(), y64, atol=1e-4)
    assert torch.allclose((O.dense_adj(src, dst, n, torch.float64) @ X.double()), y64, atol=1e-10)
    yc = c_spmm.spmm_f32(rowptr.numpy(), col.numpy(), X.numpy())
    assert np.allclose(yc, y64.numpy(), atol=1e-4)
    yc64 = c_spmm.spmm_f64acc(rowptr.numpy(), col.numpy(), X.numpy())
    assert np.allclose(yc64, y64.numpy(), atol=1e-12)


def test_batch_is_block_diagonal():
    g1 = (torch.tensor([0, 1]), torch.tensor([1, 0]), 2)
    g2 = (torch.tensor([0, 2]), torch.tensor([1, 1]), 3)
    s, d, n = O.batch_graphs([g1, g2])
    assert n == 5 and s.tolist() == [0, 1, 2, 4] and d.tolist() == [1, 0, 3, 3]
    A = O.dense_adj(s, d, n)
    assert A[:2, 2:].abs().sum() == 0 and A[2:, :2].abs().sum() == 0


def test_pos_weight_forms():
    adj = torch.zeros(10, 10)
    adj[0, 1] = adj[1, 0] = 1
    adj[2, 3] = 2
    a = O.pos_weight_inductive(adj)
    b = O.pos_weight_transductive(adj)
    assert a.dim() == 0 and b.shape == (1,)
    assert abs(float(a) - 24.0) < 1e-6 and abs(float(b) - 24.0) < 1e-6


def test_golden_vectors_reproduce():
    z = np.load(GOLDEN)
    rowptr, col = O.coo_to_csr(torch.from_numpy(z["src"]), torch.from_numpy(z["dst"]), int(z["n"]))
    assert np.array_equal(rowptr.numpy(), z["rowptr"]) and np.array_equal(col.numpy(), z["col"])
    X = torch.from_numpy(z["X"])
    assert np.allclose(O.spmm_sum(rowptr, col, X.double()).numpy(), z["spmm64"], atol=1e-12)
    weights = [(torch.from_numpy(z[f"W{i}"]), torch.from_numpy(z[f"b{i}"])) for i in range(2)]
    loss, emb, grads = O.train_step(rowptr, col, X, weights, torch.from_numpy(z["mask"]), dtype=torch.float64)
    assert abs(float(loss) - float(z["loss64"])) < 1e-10
    assert np.allclose(emb.numpy(), z["z64"], atol=1e-10)
    assert np.allclose(grads[0][0].numpy(), z["gW0_64"], atol=1e-10)
    # fp32 oracle stays within the parity tolerance of the fp64 one
    assert abs(float(z["loss32"]) - float(z["loss64"])) < 1e-5 * abs(float(z["loss64"]))


def test_train_loss_decreases_on_zinc_like_batch():
    # pin (6), qualitative: the epoch-mean loss of zinc250k.png starts ~1.1 and falls below 1.0;
    # a fresh model starts higher and must descend towards that band
    from gae_dgl_b200.synthetic import zinc_like_dataset
    ds = zinc_like_dataset(64, seed=0)
    graphs = [(*(g.edges()), g.number_of_nodes()) for g in ds]
    s, d, n = O.batch_graphs(graphs)
    rowptr, col = O.coo_to_csr(s, d, n)
    X = torch.cat([g.ndata["h"] for g in ds])
    torch.manual_seed(0)
    model = O.OracleGAE(39, [32, 16])
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    adj = O.dense_adj(s, d, n)
    pw = O.pos_weight_inductive(adj)
    losses = []
    for _ in range(150):
        loss = O.bce_loss(model(rowptr, col, X), adj, pw)
        opt.zero_grad(); loss.backward(); opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0] and losses[-1] < 1.3


# ------------------------------------------------------------------------------------------
# Pins against outputs of the reference's own code (tests/golden/make_golden_reference.py ran
# /root/reference/gae_dgl/gae.py and train_inductive.py's Trainer over a DGL stand-in)
# ------------------------------------------------------------------------------------------

from tests import _ref_fixture as RF  # noqa: E402


def _rel(a, ref):
    a, ref = a.detach().double(), ref.detach().double()
    return float((a - ref).abs().max() / max(float(ref.abs().max()), 1.0))


@pytest.mark.parametrize("tag", RF.CASES)
def test_oracle_matches_reference_run_indexing(tag):
    """Bit-exact integer work: dense adjacency (orientation A[dst, src], duplicates summed),
    in-degrees, dgl.batch offsets and the fp32 pos_weight expression."""
    c = RF.load_case(tag)
    s, d, n = O.batch_graphs([(m[0], m[1], m[2]) for m in c.members])
    assert n == c.n
    assert torch.equal(O.dense_adj(s, d, n), c.adj)
    rowptr, col = O.coo_to_csr(s, d, n)
    assert torch.equal(O.dense_adj_from_csr(rowptr, col), c.adj)
    assert torch.equal(O.in_degrees(rowptr), c.in_deg)
    assert float(O.pos_weight_inductive(c.adj)) == c.pos_weight


@pytest.mark.parametrize("tag", RF.CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_matches_reference_run_forward_and_steps(tag, dtype):
    """gae.py forward / encode and every Trainer.iteration of the reference run: logits,
    embeddings, loss and gradients per step (each step restarted from the reference's own
    weights), then the whole Adam trajectory."""
    c = RF.load_case(tag)
    s, d, n = O.batch_graphs([(m[0], m[1], m[2]) for m in c.members])
    rowptr, col = O.coo_to_csr(s, d, n)
    L = len(c.hidden)
    w0 = [(W.to(dtype), b.to(dtype)) for W, b in RF.weights_of(c.init, L)]
    emb = O.encode(rowptr, col, c.X.to(dtype), w0)
    assert _rel(emb, c.emb) < 1e-5 and _rel(emb, c.encode) < 1e-5
    assert _rel(O.decoder_logits(emb, c.masks[0], 0.1), c.logits) < 1e-5
    for step in range(len(c.losses)):
        start = c.init if step == 0 else c.after[step - 1]
        loss, _, grads = O.train_step(rowptr, col, c.X, RF.weights_of(start, L), c.masks[step], p=0.1, dtype=dtype)
        assert abs(float(loss) - c.losses[step]) < 1e-5 * abs(c.losses[step]), (tag, step)
        for i, (gW, gb) in enumerate(grads):
            rW = c.grads[step][f"layers.{i}.apply_mod.linear.weight"]
            rb = c.grads[step][f"layers.{i}.apply_mod.linear.bias"]
            assert float((gW.double() - rW.double()).abs().max()) < 2e-5 * float(rW.abs().max()), (tag, step, i)
            assert float((gb.double() - rb.double()).abs().max()) < 2e-5 * max(float(rb.abs().max()), 1e-30), (tag, step, i)
    # the Adam trajectory (train_inductive.py:40,50-52) and the evaluation call (:100-105)
    model = O.OracleGAE(c.in_dim, c.hidden)
    model.load_state_dict(c.init)            # same state_dict keys as the reference module
    opt = torch.optim.Adam(model.parameters(), lr=c.lr)
    pw = O.pos_weight_inductive(c.adj)
    for step in range(len(c.losses)):
        loss = O.bce_loss(model(rowptr, col, c.X, c.masks[step]), c.adj, pw)
        opt.zero_grad()
        loss.backward()
        opt.step()
        assert abs(float(loss) - c.losses[step]) < 2e-5 * abs(c.losses[step])
    sd = model.state_dict()
    for k, ref in c.after[-1].items():
        big = c.grads[0][k].abs() > 1e-3 * c.grads[0][k].abs().max()       # Adam's first step is sign(g)
        assert float((sd[k] - ref)[big].abs().max()) < 1e-5 + 1e-3 * c.lr, (tag, k)
    with torch.no_grad():
        ev = O.bce_loss(model(rowptr, col, c.X, c.mask_eval), c.adj, pw)
    assert abs(float(ev) - c.loss_eval) < 5e-5 * abs(c.loss_eval)


@pytest.mark.parametrize("tag", RF.CASES)
def test_module_init_is_bit_identical_to_the_reference_run(tag):
    """`torch.manual_seed(s); GAE(in_dim, hidden)` draws the same initial weights as the reference's
    constructor did in the fixture run (including the first-layer set the reference draws and discards,
    gae.py:35 vs :37/:45)."""
    import gae_dgl_b200 as G
    c = RF.load_case(tag)
    seed = {"A": 1, "B": 2, "C": 3, "D": 4}[tag]              # make_golden_reference.py: run_case(..., seed=)
    torch.manual_seed(seed)
    model = G.GAE(c.in_dim, c.hidden)
    sd = model.state_dict()
    assert list(sd.keys()) == list(c.init.keys())
    for k, ref in c.init.items():
        assert torch.equal(sd[k], ref), (tag, k)


@pytest.mark.parametrize("tag", RF.CASES)
def test_module_glue_matches_reference_run_with_cpu_test_doubles(tag, monkeypatch):
    """The Python glue of gae.py (frame handling, layer order, activations, the 'h' write-back and pop)
    executed on CPU with test doubles in place of the three CUDA ops, against the reference run."""
    import gae_dgl_b200 as G
    from gae_dgl_b200 import ops
    c = RF.load_case(tag)
    members = []
    for s, d, n, X in c.members:
        g = G.DGLGraph()
        g.add_nodes(n)
        g.add_edges(s.tolist(), d.tolist())
        g.ndata["h"] = X.clone()
        members.append(g)
    bg = G.batch(members) if len(members) > 1 else members[0]

    def spmm_double(x, graph):
        return O.spmm_sum(graph.csr().rowptr, graph.csr().col, x)

    def linear_double(y, W, b, act):
        out = F.linear(y, W, b)
        return F.relu(out) if act == ops.ACT_RELU else out

    def logits_double(z, p, mask, rng_state):
        return O.decoder_logits(z, c.masks[0] if mask is None else mask, p)

    monkeypatch.setattr(ops.SpMMFunction, "apply", staticmethod(spmm_double))
    monkeypatch.setattr(ops.LinearActFunction, "apply", staticmethod(linear_double))
    monkeypatch.setattr(ops.DecoderLogitsFunction, "apply", staticmethod(logits_double))
    model = G.GAE(c.in_dim, c.hidden)
    model.load_state_dict(c.init)
    with torch.no_grad():
        z = model.encode(bg)
        assert "h" not in bg.ndata                      # gae.py:30 pop, no write-back in encode (:57-61)
        assert _rel(z, c.encode) < 1e-5
        bg.ndata["h"] = c.X.clone()
        logits = model.forward(bg)
        assert _rel(logits, c.logits) < 1e-5
        assert _rel(bg.ndata["h"], c.emb) < 1e-5        # gae.py:53 write-back
