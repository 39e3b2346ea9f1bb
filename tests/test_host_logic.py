"""CPU: host logic, the graph container, and that the C-ABI library loads and exports every
symbol include/gae_b200.h declares (no compute calls here -- there is no GPU)."""
import ctypes
import os
import pickle
import re
import types

import numpy as np
import pytest
import torch

import gae_dgl_b200 as G
from gae_dgl_b200 import _lib, ops, synthetic
from gae_dgl_b200.graph import coo_to_csr_torch
from oracle import gae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gae_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/gae_b200.h but not exported"
    assert sorted(_lib.SIGNATURES.keys()) == names, "ctypes table and header disagree"
    assert "sm_100a" in _lib.version()


def test_sass_is_sm100a_only(lib_built):
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", lib_built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_error_codes_and_tuning(lib_built):
    lib = _lib.load()
    assert lib.gae_set_tuning(b"no_such_knob", 1) == -1
    assert b"no_such_knob" in lib.gae_last_error_string()
    _lib.set_tuning("spmm_unroll", 4)
    assert _lib.get_tuning("spmm_unroll") == 4
    _lib.set_tuning("spmm_unroll", 8)
    with pytest.raises(_lib.GaeError):
        _lib.check(lib.gae_hub_plan_count_host(None, 0, 0, None, None), "plan")


def test_ops_refuse_cpu_tensors(lib_built):
    with pytest.raises(G.GaeError):
        ops.spmm(torch.zeros(2, dtype=torch.int64), torch.zeros(0, dtype=torch.int32), torch.zeros(1, 4))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GaeError, match="no CPU fallback"):
        _lib.load()


def test_hub_plan_host_matches_numpy(lib_built):
    rng = np.random.default_rng(0)
    deg = rng.integers(0, 40, size=500)
    deg[[3, 77, 400]] = [1000, 513, 512]
    rowptr = np.zeros(501, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    plan = ops.build_hub_plan(torch.from_numpy(rowptr), seg_len=512)
    assert plan.n_long == 2 and plan.long_row[:2].tolist() == [3, 77]
    assert plan.long_seg_ptr.tolist() == [0, 2, 4] and plan.seg_row.tolist() == [0, 0, 1, 1]
    plan = ops.build_hub_plan(torch.from_numpy(rowptr), seg_len=16)
    long = np.flatnonzero(deg > 16)
    assert plan.long_row.tolist() == long.tolist()
    assert plan.n_seg == int(np.sum((deg[long] + 15) // 16))


def test_row_bins_host_matches_numpy(lib_built):
    rng = np.random.default_rng(3)
    deg = rng.integers(0, 12, size=2000)
    deg[[5, 900]] = [300, 257]
    rowptr = np.zeros(2001, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    plan = ops.build_hub_plan(torch.from_numpy(rowptr), seg_len=256, bins=True, sort_mid=False)
    empty, short, mid = (t.numpy() for t in plan.bins)
    assert empty[:plan.struct.n_empty].tolist() == np.flatnonzero(deg == 0).tolist()
    assert short[:plan.struct.n_short].tolist() == np.flatnonzero((deg >= 1) & (deg <= 4)).tolist()
    want_mid = np.flatnonzero((deg > 4) & (deg <= 256))
    assert mid[:plan.struct.n_mid].tolist() == want_mid.tolist()
    # default: the same rows ordered by descending in-degree, ties by row id (stable)
    plan_s = ops.build_hub_plan(torch.from_numpy(rowptr), seg_len=256, bins=True)
    mid_s = plan_s.bins[2].numpy()[:plan_s.struct.n_mid]
    assert mid_s.tolist() == want_mid[np.argsort(-deg[want_mid], kind="stable")].tolist()
    assert plan.long_row[:plan.n_long].tolist() == [5, 900]
    assert ops.build_hub_plan(torch.from_numpy(rowptr), seg_len=256).bins is None     # small graph: single pass


def test_graph_container_matches_oracle_indexing():
    rng = np.random.default_rng(1)
    n, e = 50, 300
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    g = G.DGLGraph()
    g.add_nodes(n)
    g.add_edges(src[:100], dst[:100])
    g.add_edges(torch.from_numpy(src[100:]), torch.from_numpy(dst[100:]))
    rp, col = O.coo_to_csr(torch.from_numpy(src), torch.from_numpy(dst), n)
    c = g.csr()
    assert torch.equal(c.rowptr, rp) and torch.equal(c.col, col)
    rt, ct = O.csr_transpose(rp, col)
    t = g.csr_t()
    assert torch.equal(t.rowptr, rt) and torch.equal(t.col, ct)
    assert torch.equal(g.in_degrees(), O.in_degrees(rp))
    assert torch.equal(g.adjacency_matrix().to_dense(), O.dense_adj(torch.from_numpy(src), torch.from_numpy(dst), n))
    assert g.number_of_edges() == e and g.number_of_nodes() == n
    r2, c2 = coo_to_csr_torch(torch.from_numpy(src), torch.from_numpy(dst), n)
    assert torch.equal(r2, rp) and torch.equal(c2, col)


def test_graph_edge_cases():
    g = G.DGLGraph()
    g.add_nodes(4)
    assert g.csr().rowptr.tolist() == [0] * 5 and g.csr().col.numel() == 0
    g.add_edges(0, [1, 2, 3])       # scalar broadcast, like DGL
    assert g.in_degrees().tolist() == [0, 1, 1, 1]
    with pytest.raises(G.GaeError):
        g.add_edges([0], [4])
    with pytest.raises(G.GaeError):
        g.update_all(lambda e: e, lambda n: n)


def test_networkx_constructor_mirrors_undirected_edges():
    import networkx as nx
    nxg = nx.Graph()
    nxg.add_nodes_from(range(4))
    nxg.add_edges_from([(0, 1), (1, 2), (3, 3)])
    g = G.DGLGraph(nxg)
    assert g.number_of_edges() == 5          # 2 x 2 mirrored + 1 self loop
    assert g.adjacency_matrix().to_dense().tolist() == [[0, 1, 0, 0], [1, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]]


def test_batch_matches_oracle_and_pickles():
    ds = synthetic.zinc_like_dataset(8, seed=3)
    bg = G.batch(ds)
    s, d, n = O.batch_graphs([(*g.edges(), g.number_of_nodes()) for g in ds])
    rp, col = O.coo_to_csr(s, d, n)
    assert bg.number_of_nodes() == n
    assert torch.equal(bg.csr().rowptr, rp) and torch.equal(bg.csr().col, col)
    assert torch.equal(bg.ndata["h"], torch.cat([g.ndata["h"] for g in ds]))
    assert bg.batch_num_nodes == [g.number_of_nodes() for g in ds]
    bg2 = pickle.loads(pickle.dumps(bg))
    assert torch.equal(bg2.csr().col, col) and torch.equal(bg2.ndata["h"], bg.ndata["h"])
    import dill
    g0 = dill.loads(dill.dumps(ds[0]))
    assert g0.number_of_edges() == ds[0].number_of_edges()


def test_zinc_like_shapes():
    ds = synthetic.zinc_like_dataset(200, seed=0)
    nodes = np.mean([g.number_of_nodes() for g in ds])
    edges = np.mean([g.number_of_edges() for g in ds])
    assert 21 < nodes < 25 and 44 < edges < 56
    for g in ds[:20]:
        assert g.ndata["h"].shape[1] == 39 and int(g.in_degrees().max()) <= 4
        a = g.adjacency_matrix().to_dense()
        assert torch.equal(a, a.t())                       # both directions (prepare_data.py:61-64)
        assert torch.equal(g.ndata["h"][:, :23].sum(1), torch.ones(g.number_of_nodes()))


def test_planetoid_like_shapes():
    g, x = synthetic.planetoid_like("cora", seed=0)
    assert g.number_of_nodes() == 2708 and g.number_of_edges() == 10556 and x.shape == (2708, 1433)
    assert torch.allclose(x.sum(1), torch.ones(2708), atol=1e-5)
    a = g.csr()
    t = g.csr_t()
    assert torch.equal(a.rowptr, t.rowptr) and torch.equal(a.col, t.col)   # symmetric


def test_rmat_is_deterministic_and_prefix_stable():
    s, d = synthetic.rmat_edges(14, 5000, seed=1)
    s2, d2 = synthetic.rmat_edges(14, 2000, seed=1, first_edge=1000, chunk=512)
    assert torch.equal(s[1000:3000], s2) and torch.equal(d[1000:3000], d2)
    assert int(s.max()) < 1 << 14 and int(s.min()) >= 0
    v = torch.arange(1 << 12)
    assert synthetic.scramble(v, 12, 5).unique().numel() == 1 << 12
    # skew survives the relabelling: some vertex collects far more than the mean in-degree
    sb, db = synthetic.rmat_edges(14, 200000, seed=1)
    deg = torch.bincount(db, minlength=1 << 14)
    assert deg.max() > 20 * deg.float().mean() and (deg == 0).float().mean() > 0.2
    x = synthetic.hashed_normal(64, 8, 2)
    assert torch.equal(x[10:20], synthetic.hashed_normal(10, 8, 2, first_row=10))


def test_module_surface_and_state_dict():
    m = G.GAE(39, [32, 16])
    assert list(m.state_dict().keys()) == list(O.OracleGAE(39, [32, 16]).state_dict().keys())
    m.load_state_dict(O.OracleGAE(39, [32, 16]).state_dict())
    assert isinstance(m.layers[0], G.GCN) and isinstance(m.layers[0].apply_mod, G.NodeApplyModule)
    assert isinstance(m.decoder, G.InnerProductDecoder) and m.decoder.dropout == 0.1
    assert m.layers[0].apply_mod.activation is torch.nn.functional.relu
    assert m.layers[1].apply_mod.activation(3) == 3
    one = G.GAE(8, [4])
    assert len(one.layers) == 1 and one.layers[0].apply_mod.activation(5) == 5
    v = G.VGAE(39, [32, 16])
    assert v.mu_head.apply_mod.linear.weight.shape == (16, 32)


def test_pos_weight_matches_reference_expression():
    g, _ = synthetic.planetoid_like("cora", seed=0)
    adj = g.adjacency_matrix().to_dense()
    assert G.pos_weight_of(g) == float(O.pos_weight_inductive(adj))
    assert G.pos_weight_of(g, transductive=True) == float(O.pos_weight_transductive(adj)[0])


def test_link_prediction_eval_utility():
    from gae_dgl_b200 import evaluate
    g, _ = synthetic.planetoid_like("cora", seed=0)
    src, dst = g.edges()
    ni, nj = evaluate.sample_non_edges(g, 500, seed=1)
    present = set((src * 2708 + dst).tolist())
    assert all((int(a) * 2708 + int(b)) not in present and a != b for a, b in zip(ni, nj))
    # embeddings that encode the graph separate edges from non-edges; random ones do not
    torch.manual_seed(0)
    z_rand = torch.randn(2708, 16)
    auc_r, ap_r = evaluate.link_prediction_metrics(z_rand, (src[:500], dst[:500]), (ni, nj))
    assert 0.4 < auc_r < 0.6
    adj = g.adjacency_matrix().to_dense()
    u, s_, _ = torch.linalg.svd(adj + torch.eye(2708))
    z_good = u[:, :64] * s_[:64].sqrt()
    auc_g, ap_g = evaluate.link_prediction_metrics(z_good, (src[:500], dst[:500]), (ni, nj))
    assert auc_g > 0.9 and ap_g > 0.9


def test_step_desc_matches_header(lib_built):
    """ctypes mirrors of the C structs have the sizes the header implies."""
    assert ctypes.sizeof(_lib.StepDesc) == 4 + 4 * 9 + 4 * 8 + 4 + 4 + 4 + 4
    assert ctypes.sizeof(_lib.HubPlanStruct) == 4 + 4 + 8 + 8 + 8 * 3 + 8 * 3 + 8 * 3 + 8
    lib = _lib.load()
    d = _lib.StepDesc()
    d.n_layers = 0
    assert lib.gae_step_ws_bytes(ctypes.byref(d), 10, None, None) == 0          # rejected, not a crash
    d.n_layers = 2
    d.dims[0], d.dims[1], d.dims[2] = 39, 32, 16
    a = lib.gae_step_ws_bytes(ctypes.byref(d), 1000, None, None)
    b = lib.gae_step_ws_bytes(ctypes.byref(d), 2000, None, None)
    assert 0 < a < b
    d.x_aggregated = 1                      # the caller brings A X: no buffer for the first aggregation
    assert lib.gae_step_ws_bytes(ctypes.byref(d), 1000, None, None) < a


def _write_planetoid(directory, name, n_train, n_all, test_ids, n_feat, n_cls, adj_lists, seed=0):
    """Writes a tiny dataset in the Planetoid on-disk layout (pickled scipy / ndarray / dict + index file)."""
    import pickle
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    n_test = len(test_ids)
    feat = lambda k: sp.csr_matrix((rng.random((k, n_feat)) < 0.3).astype(np.float32))  # noqa: E731
    lab = lambda k: np.eye(n_cls, dtype=np.int32)[rng.integers(0, n_cls, k)]  # noqa: E731
    parts = {"x": feat(n_train), "y": lab(n_train), "allx": feat(n_all), "ally": lab(n_all),
             "tx": feat(n_test), "ty": lab(n_test), "graph": adj_lists}
    for k, v in parts.items():
        with open(os.path.join(directory, f"ind.{name}.{k}"), "wb") as f:
            pickle.dump(v, f)
    with open(os.path.join(directory, f"ind.{name}.test.index"), "w") as f:
        f.write("\n".join(str(i) for i in test_ids) + "\n")
    return parts


def test_planetoid_loader(tmp_path):
    """data.load_data reads the files dgl.data.load_data would download (train_transductive.py:37-45):
    test rows restored to their node ids, Citeseer-style gaps -> zero rows, row-normalised features,
    undirected simple graph -> both directions in the DGLGraph."""
    import argparse
    from gae_dgl_b200 import data as D
    # 6 labelled/unlabelled nodes (ids 0..5), test ids {8, 6, 9}: id 7 is missing (isolated node)
    adj = {0: [1, 2], 1: [0], 2: [0, 3, 3], 3: [2], 4: [4], 5: [], 6: [0], 8: [9], 9: [8]}
    parts = _write_planetoid(str(tmp_path), "citeseer", 3, 6, [8, 6, 9], 7, 3, adj)
    parser = argparse.ArgumentParser()
    D.register_data_args(parser)
    args = parser.parse_args(["--dataset", "citeseer", "--data_dir", str(tmp_path)])
    data = D.load_data(args)
    assert data.features.shape == (10, 7) and data.features.dtype == np.float32
    tx = np.asarray(parts["tx"].todense())
    raw = np.zeros((10, 7), dtype=np.float32)
    raw[:6] = np.asarray(parts["allx"].todense())
    raw[8], raw[6], raw[9] = tx[0], tx[1], tx[2]                  # test.index order
    sums = raw.sum(1, keepdims=True)
    expect = np.divide(raw, sums, out=np.zeros_like(raw), where=sums != 0)
    assert np.allclose(data.features, expect, atol=1e-7)
    assert float(np.abs(data.features[7]).sum()) == 0.0           # the gap is an all-zero row
    assert data.labels.shape == (10,) and data.num_labels == 3
    g = G.DGLGraph(data.graph)                                    # train_transductive.py:45
    assert g.number_of_nodes() == 10
    s, d = g.edges()
    got = sorted(zip(s.tolist(), d.tolist()))
    und = {(0, 1), (0, 2), (2, 3), (0, 6), (8, 9)}                # duplicates collapsed
    want = sorted([(a, b) for a, b in und] + [(b, a) for a, b in und] + [(4, 4)])   # self loop once
    assert got == want
    assert D.find_planetoid("pubmed", str(tmp_path)) is None
    with pytest.raises(FileNotFoundError):
        D.load_data(parser.parse_args(["--dataset", "pubmed", "--data_dir", str(tmp_path)]))
    # the trainer's loader picks the files up instead of the synthetic stand-in
    from gae_dgl_b200 import train_transductive as TT
    a = TT.build_parser().parse_args(["--dataset", "citeseer", "--data_dir", str(tmp_path)])
    feats, g2 = TT.load_data(a)
    assert feats.shape == (10, 7) and g2.number_of_edges() == len(want)


def test_pos_weight_matches_the_reference_expressions_bit_for_bit():
    """gae.pos_weight_of evaluates train_inductive.py:46 / train_transductive.py:60 with numpy float32
    scalars; the reference evaluates them with 0-dim fp32 tensors (the transductive one through
    Tensor.__rtruediv__, i.e. reciprocal * scalar).  Same IEEE operations -> identical floats."""
    import random
    random.seed(0)
    for _ in range(3000):
        n = random.randint(2, 40000)
        e = random.randint(1, min(n * n - 1, 3_000_000))
        g = types.SimpleNamespace(number_of_nodes=lambda n=n: n, number_of_edges=lambda e=e: e)
        adj_sum = torch.tensor(float(e), dtype=torch.float32)          # == adj.sum() of a 0/1(/2..) fp32 matrix
        assert G.pos_weight_of(g) == float((n * n - adj_sum) / adj_sum)
        assert G.pos_weight_of(g, transductive=True) == float(torch.Tensor([float(n * n - adj_sum) / adj_sum])[0])


def test_mol_dataset_surface():
    """gae_dgl/dataset.py:3-12 surface (len, integer indexing, .graphs) + the additions, and it feeds a
    DataLoader with a collate_fn exactly as train_inductive.py:84 does."""
    from torch.utils.data import DataLoader
    graphs = synthetic.zinc_like_dataset(10, seed=3)
    ds = G.MolDataset(graphs)
    assert len(ds) == 10 and ds[3] is graphs[3] and ds.graphs[7] is graphs[7]
    assert ds[2:4] == graphs[2:4] and ds[np.array([1, 5])] == [graphs[1], graphs[5]]
    assert ds.num_nodes.tolist() == [g.number_of_nodes() for g in graphs]
    assert ds.num_edges.tolist() == [g.number_of_edges() for g in graphs]
    assert [g for g in ds] == graphs
    seen = []
    for batch in DataLoader(ds, batch_size=4, shuffle=False, collate_fn=lambda samples: samples):
        assert all(isinstance(g, G.DGLGraph) for g in batch)
        seen += batch
    assert seen == graphs


def test_native_adam_formula_matches_torch_optim():
    """The update gae_adam_step_f32 implements (csrc/misc.cu: lerp form of m, bias corrections from the
    step count in double, p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)) tracks torch.optim.Adam."""
    torch.manual_seed(0)
    p_ref = torch.randn(50, 7, requires_grad=True)
    opt = torch.optim.Adam([p_ref], lr=1e-2)
    p = p_ref.detach().clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    b1, b2, eps, lr = 0.9, 0.999, 1e-8, 1e-2
    for step in range(1, 8):
        g = torch.randn(50, 7)
        p_ref.grad = g.clone()
        opt.step()
        m = g + b1 * (m - g)
        v = b2 * v + (1.0 - b2) * g * g
        step_size = np.float32(lr / (1.0 - b1 ** step))
        inv_bc2 = np.float32(1.0 / np.sqrt(1.0 - b2 ** step))
        p = p - float(step_size) * (m / (v.sqrt() * float(inv_bc2) + eps))
        assert float((p - p_ref.detach()).abs().max()) < 2e-6 * max(1.0, float(p_ref.abs().max())), step
    st = opt.state[p_ref]
    assert float((m - st["exp_avg"]).abs().max()) < 1e-6 and float((v - st["exp_avg_sq"]).abs().max()) < 1e-6


# ---- halo planning helpers of the C ABI (host side; the kernels are covered by the -m gpu / N>1 runs) ------------

def test_halo_plan_host_matches_index_algebra(lib_built):
    """gae_halo_plan_count/fill_host against the definition: halo = sorted unique remote sources, columns
    remapped to [local | halo], per-owner counts from the block bounds."""
    from gae_dgl_b200 import parallel
    rng = np.random.default_rng(0)
    n, world = 1000, 4
    bounds = parallel.block_bounds(n, world)
    for rank in range(world):
        lo, hi = bounds[rank], bounds[rank + 1]
        src = rng.integers(0, n, size=5000).astype(np.int64)
        halo, col, recv = parallel.halo_plan_host(src, bounds, rank)
        remote = (src < lo) | (src >= hi)
        want_halo = np.unique(src[remote])
        assert np.array_equal(halo, want_halo)
        want_col = np.where(remote, (hi - lo) + np.searchsorted(want_halo, src), src - lo)
        assert np.array_equal(col, want_col.astype(np.int32))
        assert recv == [int(((want_halo >= bounds[q]) & (want_halo < bounds[q + 1])).sum()) for q in range(world)]
        assert recv[rank] == 0
    # no edges / no remote sources
    halo, col, recv = parallel.halo_plan_host(np.zeros(0, dtype=np.int64), bounds, 1)
    assert halo.size == 0 and col.size == 0 and recv == [0] * world
    halo, col, recv = parallel.halo_plan_host(np.arange(bounds[1], bounds[2], dtype=np.int64), bounds, 1)
    assert halo.size == 0 and np.array_equal(col, np.arange(bounds[2] - bounds[1], dtype=np.int32))
    # out-of-range source id is rejected
    with pytest.raises(_lib.GaeError):
        parallel.halo_plan_host(np.asarray([n], dtype=np.int64), bounds, 0)


def test_halo_stage_tags_and_push_lists_host(lib_built):
    lib = _lib.load()
    P = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    # 4 local rows, 3 halo rows (columns 4, 5, 6); two row blocks [0,2) and [2,4)
    rowptr = np.asarray([0, 2, 3, 5, 7], dtype=np.int64)
    col = np.asarray([0, 5, 1, 4, 5, 2, 6], dtype=np.int32)
    tags = np.full(3, 9, dtype=np.int32)
    rb = np.asarray([0, 2, 4], dtype=np.int64)
    _lib.check(lib.gae_halo_stage_tags_host(P(rowptr), P(col), 4, 3, P(rb), 2, P(tags)), "tags")
    assert tags.tolist() == [1, 0, 1]
    # an unreferenced halo row is an error
    assert lib.gae_halo_stage_tags_host(P(rowptr), P(col), 4, 4, P(rb), 2, P(np.zeros(4, dtype=np.int32))) == -1
    # push lists: 3 peers, I am rank 1; peer 0 asks rows [7, 8, 9] (stages 0, 1, 0), peer 2 asks [3, 4] (stages 0, 0)
    send_idx = np.asarray([7, 8, 9, 3, 4], dtype=np.int64)
    stage = np.asarray([0, 1, 0, 0, 0], dtype=np.int32)
    counts = np.asarray([3, 0, 2], dtype=np.int64)
    base = np.asarray([100, 0, 200], dtype=np.int64)
    o_src, o_peer, o_dst = np.zeros(5, dtype=np.int64), np.zeros(5, dtype=np.int32), np.zeros(5, dtype=np.int64)
    sptr = np.zeros(3, dtype=np.int64)
    _lib.check(lib.gae_halo_push_lists_host(P(send_idx), P(stage), None, P(counts), P(base), 3, 2, P(o_src), P(o_peer),
                                            P(o_dst), P(sptr)), "push lists")
    assert sptr.tolist() == [0, 4, 5]
    # stage 0 interleaves the peers (k-th entry of each), stage 1 follows; dst = base[peer] + position in the request list
    assert o_src.tolist() == [7, 3, 9, 4, 8]
    assert o_peer.tolist() == [0, 2, 0, 2, 0]
    assert o_dst.tolist() == [100, 200, 102, 201, 101]
    # explicit destinations (the consumer chose its own halo layout)
    dstpos = np.asarray([50, 51, 52, 60, 61], dtype=np.int64)
    _lib.check(lib.gae_halo_push_lists_host(P(send_idx), P(stage), P(dstpos), P(counts), None, 3, 2, P(o_src), P(o_peer),
                                            P(o_dst), P(sptr)), "push lists")
    assert o_src.tolist() == [7, 3, 9, 4, 8] and o_dst.tolist() == [50, 60, 52, 61, 51]
    assert lib.gae_halo_push_lists_host(P(send_idx), P(np.asarray([0, 2, 0, 0, 0], dtype=np.int32)), None, P(counts),
                                        P(base), 3, 2, P(o_src), P(o_peer), P(o_dst), P(sptr)) == -1
    # descriptor validation happens before any launch (no GPU needed)
    ex = _lib.HaloExchangeStruct()
    assert lib.gae_halo_push_f32(ctypes.byref(ex), 1, None) == -1
    assert lib.gae_halo_wait_f32(None, 0, 1, None) == -1


def test_fp16_pair_split_scheme_of_the_tcgen05_decoder():
    """The numeric scheme of csrc/decoder_tc16.cu restated in numpy (no GPU): operands scaled by the power of two that
    puts max|z| in [2^13, 2^14), split into fp16 hi + fp16 lo, products hi*hi + lo*hi + hi*lo accumulated exactly --
    the logits must come out within ~2^-21 of the fp64 product for embeddings of any magnitude; sigma's hi (truncated to
    11 significand bits) + lo (fp16 of the remainder) must reproduce sigma to ~1e-7 absolute."""
    rng = np.random.default_rng(0)
    for scale in (1e-6, 1e-3, 1.0, 40.0, 3e4):
        z = (rng.standard_normal((192, 16)) * scale).astype(np.float32)
        sh = int(np.floor(np.log2(np.abs(z).max()))) - 13               # the kernel: exponent field - 127 - 13
        zs = z * np.float32(2.0 ** -sh)
        assert 2.0 ** 13 <= np.abs(zs).max() < 2.0 ** 14
        hi = zs.astype(np.float16)
        lo = (zs - hi.astype(np.float32)).astype(np.float16)
        assert np.isfinite(hi.astype(np.float32)).all()
        H, L = hi.astype(np.float64), lo.astype(np.float64)
        S = (H @ H.T + L @ H.T + H @ L.T) * 4.0 ** sh
        ref = z.astype(np.float64) @ z.astype(np.float64).T
        assert np.abs(S - ref).max() / np.abs(ref).max() < 2e-6, scale
    sg = (1.0 / (1.0 + np.exp(-rng.standard_normal(100000) * 6.0))).astype(np.float32)
    hb = (sg.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
    lo = (sg - hb).astype(np.float16)
    back = hb.astype(np.float16).astype(np.float64) + lo.astype(np.float64)
    assert np.abs(back - sg.astype(np.float64)).max() < 1.5e-7
