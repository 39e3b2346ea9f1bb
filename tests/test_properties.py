"""CPU: hypothesis property tests of the integer / indexing host logic."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import gae_dgl_b200 as G
from gae_dgl_b200 import ops
from gae_dgl_b200.graph import coo_to_csr_numpy, coo_to_csr_torch
from oracle import gae_oracle as O


@st.composite
def edge_lists(draw, max_n=40, max_e=200):
    n = draw(st.integers(1, max_n))
    e = draw(st.integers(0, max_e))
    src = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    dst = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    return n, np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)


@settings(max_examples=60, deadline=None)
@given(edge_lists())
def test_csr_builders_agree_with_oracle(g):
    n, src, dst = g
    rp, col = O.coo_to_csr(torch.from_numpy(src), torch.from_numpy(dst), n)
    rn, cn = coo_to_csr_numpy(src, dst, n)
    assert np.array_equal(rn, rp.numpy()) and np.array_equal(cn, col.numpy())
    rt, ct = coo_to_csr_torch(torch.from_numpy(src), torch.from_numpy(dst), n)
    assert torch.equal(rt, rp) and torch.equal(ct, col)
    # structural invariants: monotone row pointers, sorted columns inside every row, multiset preserved
    assert rn[0] == 0 and rn[-1] == src.size and np.all(np.diff(rn) >= 0)
    for v in range(n):
        row = cn[rn[v]:rn[v + 1]]
        assert np.all(np.diff(row) >= 0)
        assert sorted(row.tolist()) == sorted(src[dst == v].tolist())
    # transpose of transpose is the identity
    t1 = O.csr_transpose(rp, col)
    t2 = O.csr_transpose(*t1)
    assert torch.equal(t2[0], rp) and torch.equal(t2[1], col)


@settings(max_examples=40, deadline=None)
@given(st.lists(edge_lists(max_n=12, max_e=30), min_size=1, max_size=6))
def test_batch_union_is_block_diagonal_and_matches_oracle(graphs):
    members = []
    for n, src, dst in graphs:
        g = G.DGLGraph()
        g.add_nodes(n)
        if src.size:
            g.add_edges(src, dst)
        g.ndata["h"] = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
        members.append(g)
    bg = G.batch(members)
    s, d, n = O.batch_graphs([(torch.from_numpy(a), torch.from_numpy(b), k) for k, a, b in graphs])
    rp, col = O.coo_to_csr(s, d, n)
    assert torch.equal(bg.csr().rowptr, rp) and torch.equal(bg.csr().col, col)
    lo, hi, pairs = bg.block_ranges()
    assert pairs == float(sum(k * k for k, _, _ in graphs))
    c = bg.csr()
    for v in range(n):          # every edge stays inside its member graph
        row = c.col[c.rowptr[v]:c.rowptr[v + 1]]
        assert bool(((row >= lo[v]) & (row < hi[v])).all())
    assert bg.ndata["h"].shape == (n, 3)


@settings(max_examples=40, deadline=None)
@given(st.lists(st.integers(0, 60), min_size=1, max_size=300), st.integers(1, 20))
def test_hub_plan_and_bins_partition_the_rows(degs, seg_len):
    deg = np.asarray(degs, dtype=np.int64)
    rowptr = np.zeros(deg.size + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    plan = ops.build_hub_plan(torch.from_numpy(rowptr), seg_len=seg_len, bins=True)
    empty, short, mid = (t.numpy() for t in plan.bins)
    ne, ns, nm = int(plan.struct.n_empty), int(plan.struct.n_short), int(plan.struct.n_mid)
    long_rows = plan.long_row.numpy()[:plan.n_long]
    rows = np.concatenate([empty[:ne], short[:ns], mid[:nm], long_rows])
    assert sorted(rows.tolist()) == list(range(deg.size))                    # exact partition of the rows
    assert np.all(deg[empty[:ne]] == 0)
    assert np.all((deg[short[:ns]] >= 1) & (deg[short[:ns]] <= min(ops.SHORT_MAX, seg_len)))
    assert np.all(deg[long_rows] > seg_len)
    # segments tile every long row exactly
    ptr = plan.long_seg_ptr.numpy()
    seg_row = plan.seg_row.numpy()[:plan.n_seg]
    for k, r in enumerate(long_rows):
        nseg = ptr[k + 1] - ptr[k]
        assert nseg == -(-deg[r] // seg_len) and np.all(seg_row[ptr[k]:ptr[k + 1]] == k)
    assert plan.n_seg == int(sum(-(-deg[r] // seg_len) for r in long_rows))
