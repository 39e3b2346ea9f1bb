"""Dataset ingestion for the transductive entry point.

The reference obtains Cora / Citeseer / Pubmed through ``dgl.data.load_data(args)`` after
``register_data_args(parser)`` (train_transductive.py:19,37-45) and reads ``data.features``
([N, F], row-normalised) and ``data.graph`` (a networkx graph handed to ``DGLGraph(...)``).
DGL downloads the Planetoid files ``ind.<name>.{x,y,tx,ty,allx,ally,graph,test.index}``; there
is no network here, so this module reads the same files from a local directory when they are
provisioned and otherwise reports that they are missing (the trainers then fall back to the
shape-faithful synthetic stand-ins of ``synthetic.py`` and say so).

Semantics kept (Planetoid / DGL citation loader):
  * feature rows = vstack(allx, tx), test rows moved back to their original node ids through
    ``test.index``; Citeseer's test indices have gaps (isolated nodes) which become all-zero rows;
  * features are row-normalised (each row divided by its sum, empty rows stay zero);
  * ``graph`` is a dict of adjacency lists -> undirected simple graph (duplicates collapse,
    self loops kept once); ``DGLGraph(nx_graph)`` then stores both directions of every edge.
"""
from __future__ import annotations

import os
import pickle
from types import SimpleNamespace
from typing import Optional

import numpy as np

PLANETOID_PARTS = ("x", "y", "tx", "ty", "allx", "ally", "graph")
DATASETS = ("cora", "citeseer", "pubmed")


def register_data_args(parser) -> None:
    """train_transductive.py:19 -- adds ``--dataset`` (plus ``--data_dir``, an addition: where the
    Planetoid files live; default $DGL_DOWNLOAD_DIR or ~/.dgl, DGL's own cache location)."""
    parser.add_argument('--dataset', type=str, default='cora', help='cora | citeseer | pubmed')
    parser.add_argument('--data_dir', type=str, default=None, help='directory holding ind.<dataset>.* files')


def default_data_dir() -> str:
    return os.environ.get("DGL_DOWNLOAD_DIR", os.path.join(os.path.expanduser("~"), ".dgl"))


def find_planetoid(name: str, data_dir: Optional[str] = None) -> Optional[str]:
    """Directory that holds ``ind.<name>.x`` (looked up in data_dir, data_dir/<name>), or None."""
    root = data_dir or default_data_dir()
    for cand in (root, os.path.join(root, name), os.path.join(root, name, "raw")):
        if os.path.exists(os.path.join(cand, f"ind.{name}.x")):
            return cand
    return None


def _read_pickle(path):
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")


def row_normalize(features) -> np.ndarray:
    """Each row divided by its sum; all-zero rows stay zero.  Accepts scipy sparse or ndarray."""
    import scipy.sparse as sp
    m = sp.csr_matrix(features, dtype=np.float64)
    sums = np.asarray(m.sum(axis=1)).reshape(-1)
    inv = np.zeros_like(sums)
    nz = sums != 0
    inv[nz] = 1.0 / sums[nz]
    return np.asarray(sp.diags(inv).dot(m).todense(), dtype=np.float32)


def load_planetoid(name: str, directory: str) -> SimpleNamespace:
    """-> namespace(features fp32 [N,F] row-normalised, graph networkx.Graph, labels int64 [N],
    num_labels, name), the attributes train_transductive.py:37-45 reads."""
    import networkx as nx
    import scipy.sparse as sp
    parts = {p: _read_pickle(os.path.join(directory, f"ind.{name}.{p}")) for p in PLANETOID_PARTS}
    with open(os.path.join(directory, f"ind.{name}.test.index")) as f:
        test_idx = np.asarray([int(line) for line in f if line.strip()], dtype=np.int64)
    tx, ty = sp.lil_matrix(parts["tx"]), np.asarray(parts["ty"])
    lo, hi = int(test_idx.min()), int(test_idx.max())
    if hi - lo + 1 != tx.shape[0]:
        # gaps in the test range (Citeseer): the missing ids are isolated nodes -> zero rows
        tx_full = sp.lil_matrix((hi - lo + 1, tx.shape[1]))
        ty_full = np.zeros((hi - lo + 1, ty.shape[1]), dtype=ty.dtype)
        order = np.sort(test_idx) - lo
        tx_full[order, :] = tx
        ty_full[order, :] = ty
        tx, ty = tx_full, ty_full
    feats = sp.vstack((sp.lil_matrix(parts["allx"]), tx)).tolil()
    labels = np.vstack((np.asarray(parts["ally"]), ty))
    sorted_idx = np.sort(test_idx)
    feats[test_idx, :] = feats[sorted_idx, :]
    labels[test_idx, :] = labels[sorted_idx, :]
    n = feats.shape[0]
    graph = nx.Graph()
    graph.add_nodes_from(range(n))
    for u, nbrs in parts["graph"].items():
        graph.add_edges_from((int(u), int(v)) for v in nbrs)
    return SimpleNamespace(name=name, features=row_normalize(feats), graph=graph,
                           labels=labels.argmax(1).astype(np.int64), num_labels=int(labels.shape[1]))


def load_data(args) -> SimpleNamespace:
    """``dgl.data.load_data(args)`` for the three citation datasets, from local files only."""
    name = args.dataset
    if name not in DATASETS:
        raise ValueError(f"unknown dataset '{name}' (expected one of {DATASETS})")
    directory = find_planetoid(name, getattr(args, "data_dir", None))
    if directory is None:
        raise FileNotFoundError(
            f"Planetoid files ind.{name}.* not found under '{getattr(args, 'data_dir', None) or default_data_dir()}' "
            "(DGL would download them; there is no network here)")
    return load_planetoid(name, directory)
