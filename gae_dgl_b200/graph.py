"""Minimal graph container exposing the DGLGraph surface the reference scripts touch.

Surface (citations relative to /root/reference/gae_dgl):
  DGLGraph() / DGLGraph(networkx_graph)            prepare_data.py:48, train_transductive.py:45
  add_nodes(n), add_edges(src, dst)                prepare_data.py:53,65
  ndata['h'] get / set / pop                       gae.py:27,30,50,53, prepare_data.py:67
  update_all(msg, reduce), apply_nodes(func=)      gae.py:28-29
  adjacency_matrix().to_dense()                    train_inductive.py:44
  in_degrees()                                     train_transductive.py:55
  to(device)                                       train_inductive.py:33
  set_n_initializer / set_e_initializer            train_inductive.py:93-94
  batch(list_of_graphs)                            train_inductive.py:34
  picklable with dill / pickle                     prepare_data.py:102-103

Indexing contract (bit-exact): CSR over destination rows, row v = sources of in-edges of v
sorted by (dst, src), duplicates kept.  adjacency_matrix() has rows = dst, cols = src
(DGL 0.4 default), so `A @ X` is exactly what update_all(copy_src, sum) computes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from ._lib import GaeError


# ---- dgl.function / dgl.init shims -----------------------------------------------------------

@dataclass(frozen=True)
class _CopySrc:
    src: str
    out: str


@dataclass(frozen=True)
class _SumReduce:
    msg: str
    out: str


class function:  # noqa: N801  (mirrors `import dgl.function as fn`)
    @staticmethod
    def copy_src(src: str, out: str) -> _CopySrc:
        return _CopySrc(src, out)

    copy_u = copy_src

    @staticmethod
    def sum(msg: str, out: str) -> _SumReduce:  # noqa: A003
        return _SumReduce(msg, out)


class init:  # noqa: N801  (mirrors dgl.init)
    @staticmethod
    def zero_initializer(shape, dtype, ctx, id_range=None):
        return torch.zeros(shape, dtype=dtype, device=ctx)


class NodeBatch:
    """What apply_nodes hands to the user function: `.data` is the node frame."""

    def __init__(self, data: Dict[str, torch.Tensor]):
        self.data = data


@dataclass
class CSR:
    rowptr: torch.Tensor          # int64 [n+1]
    col: torch.Tensor             # int32 [E]
    plan: Optional[ops.HubPlan]   # hub-row plan (CUDA only)


def _to_index_array(x) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy().astype(np.int64, copy=False).reshape(-1)
    if isinstance(x, (int, np.integer)):
        return np.asarray([x], dtype=np.int64)
    return np.asarray(list(x) if not isinstance(x, np.ndarray) else x, dtype=np.int64).reshape(-1)


def coo_to_csr_numpy(src: np.ndarray, dst: np.ndarray, n: int):
    """Host CSR build: stable sort by (dst, src).  Integer-only; bit-exact by construction."""
    if src.size == 0:
        return np.zeros(n + 1, dtype=np.int64), np.zeros(0, dtype=np.int32)
    if src.min() < 0 or dst.min() < 0 or src.max() >= n or dst.max() >= n:
        raise GaeError("edge endpoint out of range")
    order = np.lexsort((src, dst))
    col = src[order].astype(np.int32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=n), out=rowptr[1:])
    return rowptr, col


def coo_to_csr_torch(src: torch.Tensor, dst: torch.Tensor, n: int, n_cols: Optional[int] = None):
    """Same on whatever device the int64 edge tensors live on (used for the 1e8-edge RMAT
    graphs, which are generated and sorted on the GPU).  `n` rows; column ids < n_cols
    (default n) -- rectangular for the [local | halo] column space of a vertex partition."""
    m = n if n_cols is None else n_cols
    if src.numel() == 0:
        return torch.zeros(n + 1, dtype=torch.int64, device=src.device), torch.zeros(0, dtype=torch.int32, device=src.device)
    key = dst * m + src
    key, _ = torch.sort(key)
    col = (key % m).to(torch.int32)
    rows = torch.div(key, m, rounding_mode="floor")
    del key
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=src.device)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    return rowptr, col


def _plan_for(rowptr: torch.Tensor, col: torch.Tensor, seg_len: int) -> Optional[ops.HubPlan]:
    """Hub-row plan (+ degree bins, + segment order) of a device CSR; None on CPU."""
    if not rowptr.is_cuda:
        return None
    plan = ops.build_hub_plan(rowptr, seg_len)
    ops.order_segments_by_source(plan, rowptr, col)
    return plan


class DGLGraph:
    """Directed multigraph with node features; edges are (src -> dst)."""

    def __init__(self, graph_data: Any = None, *, seg_len: int = ops.DEFAULT_SEG_LEN):
        self._n = 0
        self._src: List[np.ndarray] = []
        self._dst: List[np.ndarray] = []
        self._device = torch.device("cpu")
        self.ndata: Dict[str, torch.Tensor] = {}
        self.edata: Dict[str, torch.Tensor] = {}
        self._seg_len = seg_len
        self._host_csr = None       # (rowptr, col) numpy
        self._host_csr_t = None
        self._dev_csr: Optional[CSR] = None
        self._dev_csr_t: Optional[CSR] = None
        self.batch_num_nodes: Optional[List[int]] = None
        if graph_data is not None:
            self._init_from(graph_data)

    # ---- construction --------------------------------------------------------------------
    def _init_from(self, data: Any) -> None:
        if hasattr(data, "number_of_nodes") and hasattr(data, "edges") and not isinstance(data, DGLGraph):
            # networkx graph (train_transductive.py:45).  Undirected graphs contribute both
            # directions, as DGL does when converting.
            nodes = list(data.nodes())
            index = {v: i for i, v in enumerate(nodes)}
            self.add_nodes(len(nodes))
            e = list(data.edges())
            if e:
                s = np.fromiter((index[u] for u, _ in e), dtype=np.int64, count=len(e))
                d = np.fromiter((index[v] for _, v in e), dtype=np.int64, count=len(e))
                if not data.is_directed():
                    loops = s == d
                    s, d = np.concatenate([s, d[~loops]]), np.concatenate([d, s[~loops]])
                self.add_edges(s, d)
        elif isinstance(data, tuple) and len(data) in (2, 3):
            src, dst = _to_index_array(data[0]), _to_index_array(data[1])
            n = int(data[2]) if len(data) == 3 else (int(max(src.max(), dst.max())) + 1 if src.size else 0)
            self.add_nodes(n)
            self.add_edges(src, dst)
        elif hasattr(data, "tocoo"):  # scipy sparse adjacency: A[i, j] != 0 -> edge i -> j (DGL convention)
            coo = data.tocoo()
            self.add_nodes(int(max(coo.shape)))
            self.add_edges(coo.row.astype(np.int64), coo.col.astype(np.int64))
        else:
            raise GaeError(f"cannot build a DGLGraph from {type(data).__name__}")

    @classmethod
    def from_csr(cls, rowptr: torch.Tensor, col: torch.Tensor, *, seg_len: int = ops.DEFAULT_SEG_LEN,
                 csr_t: Optional[tuple] = None) -> "DGLGraph":
        """Adopt an existing device CSR (and optionally its transpose) without a host round trip
        -- used for the large synthetic graphs."""
        g = cls(seg_len=seg_len)
        g._n = rowptr.numel() - 1
        g._device = rowptr.device
        g._frozen_edges = int(col.numel())
        g._dev_csr = CSR(rowptr, col, _plan_for(rowptr, col, seg_len))
        if csr_t is not None:
            rt, ct = csr_t
            g._dev_csr_t = CSR(rt, ct, _plan_for(rt, ct, seg_len))
        return g

    def add_nodes(self, num: int) -> None:
        self._n += int(num)
        self._invalidate()

    def add_edges(self, u, v) -> None:
        s, d = _to_index_array(u), _to_index_array(v)
        if s.size == 1 and d.size > 1:
            s = np.repeat(s, d.size)
        if d.size == 1 and s.size > 1:
            d = np.repeat(d, s.size)
        if s.size != d.size:
            raise GaeError("add_edges: src and dst lengths differ")
        if s.size and (s.min() < 0 or d.min() < 0 or s.max() >= self._n or d.max() >= self._n):
            raise GaeError("add_edges: node id out of range")
        self._src.append(s)
        self._dst.append(d)
        self._invalidate()

    add_edge = add_edges

    def _invalidate(self) -> None:
        self._host_csr = self._host_csr_t = None
        self._dev_csr = self._dev_csr_t = None

    # ---- queries ---------------------------------------------------------------------------
    def number_of_nodes(self) -> int:
        return self._n

    def number_of_edges(self) -> int:
        if getattr(self, "_frozen_edges", None) is not None and not self._src:
            return self._frozen_edges
        return int(sum(a.size for a in self._src))

    def __len__(self) -> int:
        return self._n

    @property
    def device(self) -> torch.device:
        return self._device

    def edges(self):
        """(src, dst) int64 CPU tensors in insertion order."""
        if not self._src:
            if self._dev_csr is not None:  # adopted CSR: expand (dst-sorted order)
                rp, col = self._dev_csr.rowptr, self._dev_csr.col
                deg = rp[1:] - rp[:-1]
                dst = torch.repeat_interleave(torch.arange(self._n, device=rp.device), deg)
                return col.to(torch.int64).cpu(), dst.cpu()
            z = torch.zeros(0, dtype=torch.int64)
            return z, z.clone()
        return torch.from_numpy(np.concatenate(self._src)), torch.from_numpy(np.concatenate(self._dst))

    def _coo_numpy(self):
        if not self._src:
            z = np.zeros(0, dtype=np.int64)
            return z, z
        if len(self._src) > 1:  # compact the append log
            self._src = [np.concatenate(self._src)]
            self._dst = [np.concatenate(self._dst)]
        return self._src[0], self._dst[0]

    def host_csr(self):
        if self._host_csr is None:
            s, d = self._coo_numpy()
            self._host_csr = coo_to_csr_numpy(s, d, self._n)
        return self._host_csr

    def host_csr_t(self):
        if self._host_csr_t is None:
            s, d = self._coo_numpy()
            self._host_csr_t = coo_to_csr_numpy(d, s, self._n)
        return self._host_csr_t

    def _make_dev(self, host) -> CSR:
        rowptr = torch.from_numpy(host[0]).to(self._device)
        col = torch.from_numpy(host[1]).to(self._device)
        return CSR(rowptr, col, _plan_for(rowptr, col, self._seg_len))

    def csr(self) -> CSR:
        """In-edge CSR (rows = dst) on the graph's device."""
        if self._dev_csr is None:
            self._dev_csr = self._make_dev(self.host_csr())
        return self._dev_csr

    def csr_t(self) -> CSR:
        """CSR of A^T (rows = src): drives the backward SpMM and the decoder's G^T term."""
        if self._dev_csr_t is None:
            if not self._src and self._dev_csr is not None:  # adopted CSR: transpose on device
                rp, col = self._dev_csr.rowptr, self._dev_csr.col
                deg = rp[1:] - rp[:-1]
                dst = torch.repeat_interleave(torch.arange(self._n, device=rp.device), deg)
                rt, ct = coo_to_csr_torch(dst, col.to(torch.int64), self._n)
                self._dev_csr_t = CSR(rt, ct, _plan_for(rt, ct, self._seg_len))
            else:
                self._dev_csr_t = self._make_dev(self.host_csr_t())
        return self._dev_csr_t

    def block_ranges(self):
        """(blk_lo int64 [N], blk_hi int64 [N], sum_k n_k^2) of the member graphs of a batched
        graph (a plain graph is one block) -- input of the per-graph decoder."""
        cached = getattr(self, "_block_ranges", None)
        if cached is not None and cached[0].device == self._device:
            return cached
        sizes = self.batch_num_nodes if self.batch_num_nodes else [self._n]
        sz = torch.tensor(sizes, dtype=torch.int64)
        hi = torch.cumsum(sz, 0)
        lo = hi - sz
        out = (torch.repeat_interleave(lo, sz).to(self._device), torch.repeat_interleave(hi, sz).to(self._device),
               float((sz.double() ** 2).sum()))
        self._block_ranges = out
        return out

    def in_degrees(self) -> torch.Tensor:
        """train_transductive.py:55 -- int64 in-degree per node."""
        c = self.csr()
        if c.rowptr.is_cuda:
            return ops.in_degrees(c.rowptr)
        return c.rowptr[1:] - c.rowptr[:-1]

    def out_degrees(self) -> torch.Tensor:
        c = self.csr_t()
        if c.rowptr.is_cuda:
            return ops.in_degrees(c.rowptr)
        return c.rowptr[1:] - c.rowptr[:-1]

    def adjacency_matrix(self, transpose: bool = False):
        """Sparse COO [N, N], rows = dst, cols = src, one 1.0 per edge; `.to_dense()` sums duplicates
        (train_inductive.py:44).  On a CUDA graph the result is a handle whose `.to_dense()` is DEFERRED
        (lazy.LazyAdjacency: the reference's `BCELoss(model.forward(g), adj, pos_weight)` then runs fused,
        without any N x N array) and which otherwise behaves as the sparse tensor."""
        from . import lazy
        if lazy.ENABLED and self.csr().rowptr.is_cuda:
            return lazy.AdjacencyHandle(self, transpose)
        return self.adjacency_matrix_sparse(transpose)

    def adjacency_matrix_sparse(self, transpose: bool = False) -> torch.Tensor:
        c = self.csr()
        deg = c.rowptr[1:] - c.rowptr[:-1]
        rows = torch.repeat_interleave(torch.arange(self._n, device=c.rowptr.device), deg)
        idx = torch.stack([rows, c.col.to(torch.int64)])
        if transpose:
            idx = idx.flip(0)
        vals = torch.ones(idx.shape[1], dtype=torch.float32, device=idx.device)
        return torch.sparse_coo_tensor(idx, vals, (self._n, self._n))

    # ---- device / frames -----------------------------------------------------------------
    def to(self, device) -> "DGLGraph":
        """Moves node/edge frames and the cached index structure; in place, returns self (the
        reference discards the return value, train_inductive.py:33)."""
        device = torch.device(device)
        if device != self._device:
            self._device = device
            if self._src or self._dev_csr is None:
                self._dev_csr = self._dev_csr_t = None
            else:  # adopted CSR: move the tensors
                c = self._dev_csr
                self._dev_csr = CSR(c.rowptr.to(device), c.col.to(device), None)
                self._dev_csr.plan = _plan_for(self._dev_csr.rowptr, self._dev_csr.col, self._seg_len)
                self._dev_csr_t = None
        for frame in (self.ndata, self.edata):
            for k in list(frame.keys()):
                frame[k] = frame[k].to(device)
        return self

    def set_n_initializer(self, initializer, field=None) -> None:  # train_inductive.py:94
        self._n_init = initializer

    def set_e_initializer(self, initializer, field=None) -> None:  # train_inductive.py:93
        self._e_init = initializer

    # ---- message passing (gae.py:28-29) ------------------------------------------------------
    def update_all(self, message_func, reduce_func, apply_node_func=None) -> None:
        if not (isinstance(message_func, _CopySrc) and isinstance(reduce_func, _SumReduce)
                and message_func.out == reduce_func.msg):
            raise GaeError("update_all supports the builtin pair fn.copy_src(...) / fn.sum(...) (gae.py:18-19)")
        x = self.ndata[message_func.src]
        if x.device != self._device:
            # features decide the device, like DGL frames do
            self.to(x.device)
        self.ndata[reduce_func.out] = ops.SpMMFunction.apply(x, self)
        if apply_node_func is not None:
            self.apply_nodes(apply_node_func)

    def apply_nodes(self, func=None) -> None:
        out = func(NodeBatch(self.ndata))
        self.ndata.update(out)

    # ---- pickling (prepare_data.py:102-103) ------------------------------------------------
    def __getstate__(self):
        src, dst = self.edges()
        return {
            "n": self._n, "src": src.numpy(), "dst": dst.numpy(), "seg_len": self._seg_len,
            "ndata": {k: v.detach().cpu() for k, v in self.ndata.items()},
            "edata": {k: v.detach().cpu() for k, v in self.edata.items()},
            "batch_num_nodes": self.batch_num_nodes,
        }

    def __setstate__(self, st):
        self.__init__(seg_len=st.get("seg_len", ops.DEFAULT_SEG_LEN))
        self._n = st["n"]
        if st["src"].size:
            self._src, self._dst = [st["src"]], [st["dst"]]
        self.ndata, self.edata = st["ndata"], st["edata"]
        self.batch_num_nodes = st.get("batch_num_nodes")

    def __repr__(self):
        return f"DGLGraph(num_nodes={self._n}, num_edges={self.number_of_edges()}, device={self._device})"


def batch(graphs: Sequence[DGLGraph], device=None) -> DGLGraph:
    """dgl.batch (train_inductive.py:34): block-diagonal disjoint union.  Node ids of graph k
    are offset by the prefix sum of node counts; node frames are concatenated.

    The union's CSR is assembled directly from the members' cached host CSRs (concatenate,
    one H2D copy, one kernel that adds the node offsets: gae_batch_offset_cols_i32) instead
    of re-sorting the edge list every step."""
    graphs = list(graphs)
    if not graphs:
        raise GaeError("batch() needs at least one graph")
    if device is None:
        device = graphs[0].device
        for g in graphs:
            for v in g.ndata.values():
                device = v.device
                break
            break
    device = torch.device(device)
    sizes = np.asarray([g.number_of_nodes() for g in graphs], dtype=np.int64)
    node_off = np.zeros(len(graphs) + 1, dtype=np.int64)
    np.cumsum(sizes, out=node_off[1:])
    bg = DGLGraph(seg_len=graphs[0]._seg_len)
    bg._n = int(node_off[-1])
    bg._device = device
    bg.batch_num_nodes = sizes.tolist()

    def union(get):
        parts = [get(g) for g in graphs]
        ecount = np.asarray([p[1].size for p in parts], dtype=np.int64)
        edge_ptr = np.zeros(len(graphs) + 1, dtype=np.int64)
        np.cumsum(ecount, out=edge_ptr[1:])
        rowptr = np.zeros(bg._n + 1, dtype=np.int64)
        pos = 1
        for k, (rp, _) in enumerate(parts):
            m = rp.size - 1
            rowptr[pos:pos + m] = rp[1:] + edge_ptr[k]
            pos += m
        col_local = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.int32)
        return rowptr, col_local, edge_ptr

    if device.type == "cuda":
        from . import _lib
        import ctypes
        lib = _lib.load()
        off_dev = torch.from_numpy(node_off).to(device)
        out = []
        for get in (DGLGraph.host_csr, DGLGraph.host_csr_t):
            rowptr, col_local, edge_ptr = union(get)
            rp_dev = torch.from_numpy(rowptr).to(device)
            col_dev = torch.from_numpy(col_local).to(device)
            ep_dev = torch.from_numpy(edge_ptr).to(device)
            rc = lib.gae_batch_offset_cols_i32(ctypes.c_void_p(col_dev.data_ptr()), ctypes.c_void_p(ep_dev.data_ptr()),
                                               ctypes.c_void_p(off_dev.data_ptr()), len(graphs), col_dev.numel(),
                                               ops._stream())
            _lib.check(rc, "gae_batch_offset_cols_i32")
            # molecular batches have max degree 4: no hub rows, skip the host plan scan
            out.append(CSR(rp_dev, col_dev, None))
        bg._dev_csr, bg._dev_csr_t = out
        bg._frozen_edges = int(out[0].col.numel())
    # keep the edge log too (cheap) so edges()/pickling/adjacency on CPU work
    srcs, dsts = [], []
    for k, g in enumerate(graphs):
        s, d = g._coo_numpy()
        srcs.append(s + node_off[k])
        dsts.append(d + node_off[k])
    bg._src = [np.concatenate(srcs)] if srcs else []
    bg._dst = [np.concatenate(dsts)] if dsts else []
    # node frames: concatenate every key present in all members
    keys = set(graphs[0].ndata.keys())
    for g in graphs[1:]:
        keys &= set(g.ndata.keys())
    for k in keys:
        bg.ndata[k] = torch.cat([g.ndata[k] for g in graphs], 0).to(device)
    return bg


class PackedGraphDataset:
    """All graphs of a dataset resident on the device as ONE packed CSR (+ its transpose + node
    features), so a mini-batch is assembled by a kernel instead of a host loop over its members
    (after fusing the decoder, `collate` on the Python main thread
    -- train_inductive.py:31-35,84 -- becomes the bottleneck of the inductive loop).

    `batch(ids)` returns the same block-diagonal DGLGraph as `batch([graphs[i] for i in ids])`,
    bit-identical indexing (tested), with ndata[feature_key] gathered on the device."""

    def __init__(self, graphs: Sequence[DGLGraph], device, feature_key: str = "h"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise GaeError("PackedGraphDataset lives on a CUDA device")
        torch.cuda.set_device(self.device)     # its batches are assembled by native kernels on the current device
        self.feature_key = feature_key
        self.n_graphs = len(graphs)
        nodes = np.asarray([g.number_of_nodes() for g in graphs], dtype=np.int64)
        self.nodes = nodes
        self.node_ptr_host = np.zeros(len(graphs) + 1, dtype=np.int64)
        np.cumsum(nodes, out=self.node_ptr_host[1:])
        self._seg_len = graphs[0]._seg_len if graphs else ops.DEFAULT_SEG_LEN

        def pack(get):
            parts = [get(g) for g in graphs]
            edges = np.asarray([p[1].size for p in parts], dtype=np.int64)
            eptr = np.zeros(len(graphs) + 1, dtype=np.int64)
            np.cumsum(edges, out=eptr[1:])
            rowptr = np.zeros(int(self.node_ptr_host[-1]) + 1, dtype=np.int64)
            pos = 1
            for k, (rp, _) in enumerate(parts):
                m = rp.size - 1
                rowptr[pos:pos + m] = rp[1:] + eptr[k]
                pos += m
            col = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, np.int32)
            return edges, torch.from_numpy(rowptr).to(self.device), torch.from_numpy(col).to(self.device)

        self.edges, self.rowptr_all, self.col_all = pack(DGLGraph.host_csr)
        self.edges_t, self.rowptr_all_t, self.col_all_t = pack(DGLGraph.host_csr_t)
        self.node_ptr = torch.from_numpy(self.node_ptr_host).to(self.device)
        feats = [g.ndata[feature_key] for g in graphs] if graphs and feature_key in graphs[0].ndata else None
        self.feat_all = ops.as_rows(torch.cat(feats, 0).to(self.device, torch.float32), "features") if feats else None

    def __len__(self):
        return self.n_graphs

    def batch(self, ids) -> DGLGraph:
        import ctypes
        from . import _lib
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        k = ids.size
        if k == 0:
            raise GaeError("batch() needs at least one graph")
        noff = np.zeros(k + 1, dtype=np.int64)
        np.cumsum(self.nodes[ids], out=noff[1:])
        n_out = int(noff[-1])
        lib = _lib.load()
        dev = self.device
        # ONE truly asynchronous H2D for all index vectors: they are staged in pinned memory
        # (a pageable source would make cudaMemcpyAsync synchronise the stream every batch)
        eoffs = []
        for edges in (self.edges, self.edges_t):
            eoff = np.zeros(k + 1, dtype=np.int64)
            np.cumsum(edges[ids], out=eoff[1:])
            eoffs.append(eoff)
        staged = torch.from_numpy(np.concatenate([ids, noff, eoffs[0], eoffs[1]])).pin_memory()
        on_dev = staged.to(dev, non_blocking=True)
        gid_dev, noff_dev = on_dev[:k], on_dev[k:2 * k + 1]
        eoff_devs = (on_dev[2 * k + 1:3 * k + 2], on_dev[3 * k + 2:4 * k + 3])
        bg = DGLGraph(seg_len=self._seg_len)
        bg._n = n_out
        bg._device = dev
        bg.batch_num_nodes = self.nodes[ids].tolist()
        p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())  # noqa: E731
        out = []
        for which, (rowptr_all, col_all, with_feat) in enumerate(((self.rowptr_all, self.col_all, True),
                                                                  (self.rowptr_all_t, self.col_all_t, False))):
            eoff, eoff_dev = eoffs[which], eoff_devs[which]
            e_out = int(eoff[-1])
            rp = torch.empty(n_out + 1, dtype=torch.int64, device=dev)
            col = torch.empty(e_out, dtype=torch.int32, device=dev)
            feat_out = None
            if with_feat and self.feat_all is not None:
                feat_out = ops.alloc_rows(n_out, self.feat_all.shape[1], dev)
            rc = lib.gae_batch_assemble(p(rowptr_all), p(col_all), p(self.node_ptr), p(gid_dev), k, p(noff_dev),
                                        p(eoff_dev), n_out, e_out, p(rp), p(col),
                                        p(self.feat_all) if feat_out is not None else None,
                                        self.feat_all.stride(0) if feat_out is not None else 0,
                                        self.feat_all.shape[1] if feat_out is not None else 0, p(feat_out),
                                        feat_out.stride(0) if feat_out is not None else 0,
                                        ops._stream())
            _lib.check(rc, "gae_batch_assemble")
            if feat_out is not None:
                bg.ndata[self.feature_key] = feat_out
            out.append(CSR(rp, col, None))
        bg._dev_csr, bg._dev_csr_t = out
        bg._frozen_edges = int(out[0].col.numel())
        return bg
