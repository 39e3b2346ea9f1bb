"""CPU: a ``graphs.pkl`` whose graphs are ``dgl.DGLGraph`` objects (prepare_data.py:102-103) is
read without DGL (gae_dgl_b200/dgl_pickle.py).

No DGL build exists here, so the writer is a throw-away look-alike ``dgl`` package with the
pickling layout of DGL 0.3 / 0.4 (GraphIndex.__getstate__ tuple, utils.Index state, FrameRef ->
Frame -> Column), created in a temp directory and used in a SUBPROCESS only; this process reads
the file with no ``dgl`` importable.  That pins the reader against the layout it documents, not
against a real DGL file (dgl_pickle.py says so in its header)."""
import subprocess
import sys
import textwrap

import pytest
import torch

import gae_dgl_b200 as G
from gae_dgl_b200 import dgl_pickle
from gae_dgl_b200._lib import GaeError

FAKE_DGL = {
    "__init__.py": """
        from .graph import DGLGraph
        from . import init
    """,
    "init.py": """
        def zero_initializer(shape, dtype, ctx, id_range=None):
            import torch
            return torch.zeros(shape, dtype=dtype)
    """,
    "utils.py": """
        import torch
        LAYOUT = "0.4"
        class Index(object):
            def __init__(self, data, dtype="int64"):
                self._slice_data = data if isinstance(data, slice) else None
                self._user_tensor_data = {} if isinstance(data, slice) else {"cpu": torch.as_tensor(data, dtype=torch.int64)}
                self._dtype = dtype
            def tousertensor(self):
                if self._slice_data is not None:
                    s = self._slice_data
                    return torch.arange(s.start, s.stop, dtype=torch.int64)
                return self._user_tensor_data["cpu"]
            def __getstate__(self):
                if LAYOUT == "0.3":
                    return self.tousertensor()
                if self._slice_data is not None:
                    return self._slice_data, self._dtype
                return self.tousertensor(), self._dtype
            def __setstate__(self, state):
                raise RuntimeError("the look-alike dgl must never be imported by the reader")
    """,
    "graph_index.py": """
        from enum import Enum
        from . import utils
        class BoolFlag(Enum):
            BOOL_UNKNOWN = -1
            BOOL_FALSE = 0
            BOOL_TRUE = 1
        class GraphIndex(object):
            def __init__(self):
                self.n = 0
                self.src = []
                self.dst = []
                self._cache = {"handle": object}
            def __getstate__(self):
                n_edges = len(self.src)
                src, dst = utils.Index(self.src), utils.Index(self.dst)
                if utils.LAYOUT == "0.3":
                    return self.n, True, src, dst                 # n_nodes, readonly, src, dst
                if utils.LAYOUT == "0.4-flag":
                    return self.n, BoolFlag.BOOL_TRUE, False, src, dst
                return self.n, True, False, src, dst              # n_nodes, multigraph, readonly, src, dst
            def __setstate__(self, state):
                raise RuntimeError("the look-alike dgl must never be imported by the reader")
    """,
    "frame.py": """
        from collections import namedtuple
        from . import utils, init
        class Scheme(namedtuple("Scheme", ["shape", "dtype"])):
            pass
        class Column(object):
            def __init__(self, data):
                self.data = data
                self.scheme = Scheme(tuple(data.shape[1:]), data.dtype)
        class Frame(object):
            def __init__(self, num_rows=0):
                self._columns = {}
                self._num_rows = num_rows
                self._initializers = {}
                self._remote_init_builder = None
                self._default_initializer = init.zero_initializer
        class FrameRef(object):
            def __init__(self, frame):
                self._frame = frame
                self._index = utils.Index(slice(0, frame._num_rows))
                self._index_data = None
    """,
    "graph.py": """
        from .graph_index import GraphIndex
        from .frame import Frame, FrameRef, Column
        class DGLBaseGraph(object):
            def __init__(self):
                self._graph = GraphIndex()
        class DGLGraph(DGLBaseGraph):
            def __init__(self):
                super().__init__()
                self._node_frame = FrameRef(Frame(0))
                self._edge_frame = FrameRef(Frame(0))
                self._msg_index = None
                self._msg_frame = FrameRef(Frame(0))
                self._message_func = None
                self._reduce_func = None
                self._apply_node_func = None
                self._apply_edge_func = None
            def add_nodes(self, n):
                self._graph.n += n
                self._node_frame = FrameRef(Frame(self._graph.n))
            def add_edges(self, u, v):
                self._graph.src += list(u)
                self._graph.dst += list(v)
                self._edge_frame = FrameRef(Frame(len(self._graph.src)))
            def set_ndata(self, name, t):
                self._node_frame._frame._columns[name] = Column(t)
            def set_edata(self, name, t):
                self._edge_frame._frame._columns[name] = Column(t)
    """,
}

WRITER = """
import sys, dill, torch
sys.path.insert(0, sys.argv[1])
import dgl, dgl.utils
dgl.utils.LAYOUT = sys.argv[3]
torch.manual_seed(0)
graphs = []
# prepare_data.py:48-67 -- nodes, both directions of every bond, 'h' features
specs = [(5, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0)]), (3, [(0, 1), (0, 1), (1, 2)]), (4, []), (1, [(0, 0)])]
for n, bonds in specs:
    g = dgl.DGLGraph()
    g.add_nodes(n)
    src, dst = [], []
    for a, b in bonds:
        src.extend([a, b]); dst.extend([b, a])
    if src:
        g.add_edges(src, dst)
    g.set_ndata('h', torch.arange(n * 39, dtype=torch.float32).reshape(n, 39) / 7)
    if src:
        g.set_edata('w', torch.arange(len(src), dtype=torch.float32))
    graphs.append(g)
with open(sys.argv[2], 'wb') as f:
    dill.dump(graphs, f)
"""

SPECS = [(5, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0)]), (3, [(0, 1), (0, 1), (1, 2)]), (4, []), (1, [(0, 0)])]


def write_with_lookalike(tmp_path, layout):
    pytest.importorskip("dill")
    pkg = tmp_path / "fake_site" / "dgl"
    pkg.mkdir(parents=True)
    for name, src in FAKE_DGL.items():
        (pkg / name).write_text(textwrap.dedent(src))
    out = tmp_path / f"graphs_{layout}.pkl"
    script = tmp_path / "writer.py"
    script.write_text(WRITER)
    r = subprocess.run([sys.executable, str(script), str(tmp_path / "fake_site"), str(out), layout],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    return str(out)


@pytest.mark.parametrize("layout", ["0.4", "0.4-flag", "0.3"])
def test_reads_dgl_written_graphs_without_dgl(tmp_path, layout):
    path = write_with_lookalike(tmp_path, layout)
    graphs = dgl_pickle.load_graph_list(path)
    assert len(graphs) == len(SPECS)
    for g, (n, bonds) in zip(graphs, SPECS):
        assert isinstance(g, G.DGLGraph)
        src = [x for a, b in bonds for x in (a, b)]
        dst = [x for a, b in bonds for x in (b, a)]
        assert g.number_of_nodes() == n and g.number_of_edges() == len(src)
        s, d = g.edges()
        assert s.tolist() == src and d.tolist() == dst                # insertion order = edge ids
        ref = torch.arange(n * 39, dtype=torch.float32).reshape(n, 39) / 7
        assert torch.equal(g.ndata['h'], ref)
        if src:
            assert torch.equal(g.edata['w'], torch.arange(len(src), dtype=torch.float32))
        # A[dst, src] with duplicates summed, the contract of train_inductive.py:44
        dense = g.adjacency_matrix_sparse().to_dense()
        want = torch.zeros(n, n)
        for a, b in zip(src, dst):
            want[b, a] += 1
        assert torch.equal(dense, want)
    # the batch the trainer collates from them (train_inductive.py:34)
    bg = G.batch(graphs)
    assert bg.number_of_nodes() == sum(n for n, _ in SPECS)
    assert bg.ndata['h'].shape == (bg.number_of_nodes(), 39)


def test_own_graphs_pass_through_and_bad_files_fail_loudly(tmp_path):
    dill = pytest.importorskip("dill")
    g = G.DGLGraph()
    g.add_nodes(3)
    g.add_edges([0, 1], [1, 2])
    g.ndata['h'] = torch.ones(3, 39)
    p = tmp_path / "own.pkl"
    with open(p, "wb") as f:
        dill.dump([g, g], f)                                          # prepare_data.py:102-103 over this package's graphs
    back = dgl_pickle.load_graph_list(str(p))
    assert len(back) == 2 and isinstance(back[0], G.DGLGraph) and back[0].number_of_edges() == 2
    assert torch.equal(back[0].ndata['h'], g.ndata['h'])

    with open(p, "wb") as f:
        dill.dump({"not": "a list"}, f)
    with pytest.raises(GaeError, match="expected a pickled list"):
        dgl_pickle.load_graph_list(str(p))

    # a dgl object that carries no GraphIndex state is refused, not guessed
    stub = dgl_pickle._stub_for("dgl.graph", "DGLGraph")()
    stub.__setstate__({"_node_frame": None})
    with pytest.raises(GaeError, match="GraphIndex"):
        dgl_pickle.graph_from_dgl_state(stub)


def test_train_inductive_loads_a_dgl_written_file(tmp_path):
    from gae_dgl_b200 import train_inductive as TI
    path = write_with_lookalike(tmp_path, "0.4")
    args = TI.build_parser().parse_args(["--data_file", path])
    graphs = TI.load_graphs(args)
    assert len(graphs) == len(SPECS) and all(isinstance(g, G.DGLGraph) for g in graphs)


@pytest.mark.gpu
def test_dgl_written_file_trains_on_the_gpu(cuda, tmp_path):
    """train_inductive.py:79-81 -> :34 -> :43-53 over a DGL-written file: the graphs read without DGL are
    collated on the device and one step matches the fp64 oracle (1e-5, the bar of tests/test_parity_gpu.py)."""
    from oracle import gae_oracle as O
    from tests.test_parity_gpu import _check_step
    graphs = dgl_pickle.load_graph_list(write_with_lookalike(tmp_path, "0.4"))
    bg = G.batch(graphs, device=cuda)
    s, d, n = O.batch_graphs([(*g.edges(), g.number_of_nodes()) for g in graphs])
    rp, col = O.coo_to_csr(s, d, n)
    assert torch.equal(bg.csr().rowptr.cpu(), rp) and torch.equal(bg.csr().col.cpu(), col)      # bit-exact indexing
    X = bg.ndata["h"].cpu() / 40.0
    torch.manual_seed(3)
    ref = O.OracleGAE(39, [32, 16])
    weights = [(l.apply_mod.linear.weight.detach(), l.apply_mod.linear.bias.detach()) for l in ref.layers]
    mask = torch.rand(n, 16) >= 0.1
    _check_step(cuda, bg, X, weights, mask)
