"""Reads a ``graphs.pkl`` written by the reference's ``prepare_data.py`` WITHOUT DGL installed
(SURVEY.md section 8f rank 4).

The reference dumps ``list[dgl.DGLGraph]`` with dill (prepare_data.py:102-103) and the trainer
reads it back with ``dill.load`` (train_inductive.py:79-81), which needs the very DGL build that
wrote the file.  Here every class / function the stream names under the ``dgl`` package is
replaced by a state-capturing stand-in, and the graph is recovered from the captured state:

  * structure: DGL 0.3/0.4 ``GraphIndex.__getstate__`` -> ``(n_nodes, [multigraph,] readonly,
    src, dst)`` with ``src`` / ``dst`` as ``utils.Index`` objects whose own state is the id
    tensor (optionally followed by a dtype string) or a ``slice``;
  * node features: ``DGLGraph._node_frame`` (FrameRef) -> ``_frame`` (Frame) -> ``_columns``
    ``{name: Column}`` -> ``Column.data`` (a torch tensor); edge features likewise from
    ``_edge_frame``.

The walk does not rely on the exact attribute nesting (it searches the captured object tree for
a ``GraphIndex`` state and for ``_columns`` dicts under the node / edge frame attributes), so
small layout differences between DGL point releases are tolerated; what cannot be interpreted
raises ``GaeError`` naming the offending element -- nothing is guessed silently.  Edge order
(hence edge ids) and duplicate edges are preserved: the graphs are multigraphs
(prepare_data.py:61-65 adds both directions of every bond).

Parity status: no DGL build is installable here, so this reader is pinned only against a
look-alike writer (tests/test_dgl_pickle.py builds a throw-away ``dgl`` package with the layout
above, pickles graphs with it in a subprocess and reads them back here); it has not seen a file
written by a real DGL.  Objects of this package's own ``DGLGraph`` class pass through untouched,
so files written by this framework load through the same call.
"""
from __future__ import annotations

import io
import pickle
from typing import Any, Dict, List, Tuple

import numpy as np
import torch

from ._lib import GaeError

_STUBS: Dict[Tuple[str, str], type] = {}


class _DGLStub:
    """Stand-in for anything the pickle stream imports from ``dgl``: keeps constructor arguments
    (``_args``) and whatever state the stream restores (``_state``: a dict for default pickling,
    the ``__getstate__`` value otherwise)."""
    _dgl_module = ""
    _dgl_name = ""

    def __new__(cls, *args, **kwargs):
        obj = object.__new__(cls)
        obj._args = args
        obj._state = None
        return obj

    def __init__(self, *args, **kwargs):
        pass

    def __setstate__(self, state):
        self._state = state

    def __repr__(self):
        return f"<dgl stand-in {self._dgl_module}.{self._dgl_name}>"


def _stub_for(module: str, name: str) -> type:
    key = (module, name)
    if key not in _STUBS:
        _STUBS[key] = type(name, (_DGLStub,), {"_dgl_module": module, "_dgl_name": name})
    return _STUBS[key]


def _unpickler_base():
    try:                       # the file was written by dill; its Unpickler also resolves dill's helpers
        import dill
        return dill.Unpickler
    except ImportError:
        return pickle.Unpickler


def _make_unpickler(f):
    base = _unpickler_base()

    class _Unpickler(base):
        def find_class(self, module, name):
            if module == "dgl" or module.startswith("dgl."):
                return _stub_for(module, name)
            return super().find_class(module, name)

    return _Unpickler(f)


# ---- interpreting the captured state --------------------------------------------------------------

def _children(obj: Any):
    """(key, child) pairs of one captured object; key is the attribute / dict key / position."""
    if isinstance(obj, _DGLStub):
        st = obj._state
        if isinstance(st, dict):
            yield from st.items()
        elif isinstance(st, (tuple, list)):
            # (dict, slots_dict) is what default pickling produces for classes with __slots__
            yield from enumerate(st)
        elif st is not None:
            yield "_state", st
        yield from (("_arg%d" % i, a) for i, a in enumerate(obj._args))
    elif isinstance(obj, dict):
        yield from obj.items()
    elif isinstance(obj, (tuple, list)):
        yield from enumerate(obj)


def _find(obj: Any, pred, _seen=None, _depth=0):
    """Depth-first search of the captured tree for the first object satisfying pred."""
    if _seen is None:
        _seen = set()
    if id(obj) in _seen or _depth > 12:
        return None
    _seen.add(id(obj))
    if pred(obj):
        return obj
    for _, child in _children(obj):
        if isinstance(child, (_DGLStub, dict, tuple, list)):
            hit = _find(child, pred, _seen, _depth + 1)
            if hit is not None:
                return hit
    return None


def _attr(obj: Any, *names: str):
    """First present attribute of a stand-in (from its state dict), else None."""
    if isinstance(obj, _DGLStub) and isinstance(obj._state, dict):
        for n in names:
            if n in obj._state:
                return obj._state[n]
    return None


def _is_stub(obj: Any, name: str) -> bool:
    return isinstance(obj, _DGLStub) and obj._dgl_name == name


def _ids_of(obj: Any, what: str) -> np.ndarray:
    """utils.Index state (or a bare tensor / array / list / slice) -> int64 numpy ids."""
    if _is_stub(obj, "Index"):
        st = obj._state
        if isinstance(st, dict):        # default pickling: whichever materialised form the Index holds
            for k in ("_user_tensor_data", "_pydata", "_slice_data"):
                held = st.get(k)
                if isinstance(held, dict):                   # {context: tensor}
                    held = next(iter(held.values()), None)
                if held is not None:
                    return _ids_of(held, what)
            raise GaeError(f"graphs.pkl: cannot read the {what} ids of a DGL Index with state keys {sorted(st)}")
        if isinstance(st, (tuple, list)) and len(st) in (1, 2) and not isinstance(st[0], (int, np.integer)):
            return _ids_of(st[0], what)          # (data, dtype) of DGL 0.4
        return _ids_of(st, what)
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().to(torch.int64).numpy()
    if isinstance(obj, slice):
        return np.arange(obj.start or 0, obj.stop, obj.step or 1, dtype=np.int64)
    if isinstance(obj, np.ndarray) or (isinstance(obj, (list, tuple)) and all(isinstance(i, (int, np.integer)) for i in obj)):
        return np.asarray(obj, dtype=np.int64).reshape(-1)
    raise GaeError(f"graphs.pkl: cannot read {what} ids from {type(obj).__name__}")


def _graph_index_state(root: Any):
    """-> (n_nodes, src ids, dst ids) from the GraphIndex stand-in under root."""
    gi = _find(root, lambda o: _is_stub(o, "GraphIndex") and isinstance(o._state, (tuple, list)))
    if gi is None:
        raise GaeError("graphs.pkl: no DGL GraphIndex state (n_nodes, ..., src, dst) found in a pickled graph")
    st = list(gi._state)
    ints = [x for x in st if isinstance(x, (int, np.integer)) and not isinstance(x, (bool, np.bool_))]
    idx = [x for x in st if _is_stub(x, "Index") or isinstance(x, (torch.Tensor, np.ndarray, slice))]
    if len(ints) < 1 or len(idx) != 2:
        raise GaeError(f"graphs.pkl: unexpected GraphIndex state layout {[type(x).__name__ for x in st]}")
    src, dst = _ids_of(idx[0], "source"), _ids_of(idx[1], "destination")
    if src.shape != dst.shape:
        raise GaeError("graphs.pkl: source and destination id arrays differ in length")
    return int(ints[0]), src, dst


def _frame_columns(root: Any, frame_attr: str, n_rows: int) -> Dict[str, torch.Tensor]:
    """{name: tensor} of the Frame reached from root.<frame_attr> (FrameRef -> Frame -> _columns)."""
    ref = _attr(root, frame_attr)
    if ref is None:
        ref = _find(root, lambda o: isinstance(o, _DGLStub) and isinstance(o._state, dict) and frame_attr in o._state)
        ref = _attr(ref, frame_attr) if ref is not None else None
    if ref is None:
        return {}
    holder = _find(ref, lambda o: isinstance(o, _DGLStub) and isinstance(o._state, dict) and "_columns" in o._state)
    if holder is None:
        return {}
    out: Dict[str, torch.Tensor] = {}
    for name, col in holder._state["_columns"].items():
        data = col if isinstance(col, torch.Tensor) else _attr(col, "data", "_data")
        if isinstance(data, np.ndarray):
            data = torch.from_numpy(data)
        if not isinstance(data, torch.Tensor):
            raise GaeError(f"graphs.pkl: column '{name}' of {frame_attr} holds {type(data).__name__}, not a tensor")
        # a FrameRef may view a subset of the frame's rows; the reference never creates such graphs
        index = _attr(ref, "_index")
        if index is not None and data.shape[0] != n_rows:
            data = data[torch.from_numpy(_ids_of(index, "frame row"))]
        if data.shape[0] != n_rows:
            raise GaeError(f"graphs.pkl: column '{name}' of {frame_attr} has {data.shape[0]} rows, the graph has {n_rows}")
        out[str(name)] = data
    return out


def graph_from_dgl_state(obj: Any):
    """One captured ``dgl.DGLGraph`` -> this package's DGLGraph (same nodes, edges in the same
    order, same ndata / edata tensors)."""
    from .graph import DGLGraph
    n, src, dst = _graph_index_state(obj)
    g = DGLGraph()
    g.add_nodes(n)
    if src.size:
        g.add_edges(src, dst)
    for k, v in _frame_columns(obj, "_node_frame", n).items():
        g.ndata[k] = v
    for k, v in _frame_columns(obj, "_edge_frame", int(src.size)).items():
        g.edata[k] = v
    return g


def _convert(obj: Any):
    if isinstance(obj, _DGLStub):
        return graph_from_dgl_state(obj)
    if isinstance(obj, list):
        return [_convert(x) for x in obj]
    if isinstance(obj, tuple):
        return tuple(_convert(x) for x in obj)
    if isinstance(obj, dict):
        return {k: _convert(v) for k, v in obj.items()}
    return obj


def loads(data: bytes):
    return _convert(_make_unpickler(io.BytesIO(data)).load())


def load(path_or_file) -> Any:
    """``dill.load`` of a graphs file with DGL graphs turned into this package's DGLGraph; the
    container shape (the reference writes a flat list) is kept."""
    if hasattr(path_or_file, "read"):
        return _convert(_make_unpickler(path_or_file).load())
    with open(path_or_file, "rb") as f:
        return _convert(_make_unpickler(f).load())


def load_graph_list(path) -> List[Any]:
    """train_inductive.py:79-81 -- the list of graphs of a ``graphs.pkl``."""
    graphs = load(path)
    if not isinstance(graphs, (list, tuple)):
        raise GaeError(f"{path}: expected a pickled list of graphs, found {type(graphs).__name__}")
    return list(graphs)
