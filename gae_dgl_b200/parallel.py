"""1-D vertex partition of the encoder SpMM across the GPUs of one box.

Rank r owns the contiguous block of destination rows [r*V/P, (r+1)*V/P), their in-edge CSR and
the feature rows of those vertices.  Column ids are remapped to [local | halo]: halo = sorted
unique remote sources, which (blocks being contiguous) are already grouped by owner.  One
exchange per SpMM moves each needed remote row exactly once:

  nccl : pack (gae_gather_rows_f32) -> all-to-all-v (NCCL grouped send/recv over NVLink)
         -> rows land directly in the halo region of the [local | halo] feature buffer
  p2p  : one-sided pull -- every rank maps its peers' feature buffers through CUDA IPC and
         gathers the rows it needs straight out of peer HBM over NVSwitch
         (gae_pull_rows_p2p_f32): no pack, no staging copy, no sender-side kernel.

The backward SpMM (dX = A^T dY) uses the same machinery on CSR(A^T), partitioned by source.
Graph generation is distributed as well: every rank draws 1/P of the R-MAT edge stream and
routes each edge to the owner of its row.

`pack_fn` / `spmm_fn` default to the CUDA kernels; the gloo CPU tests inject test doubles for
them so the planning / exchange logic runs without a GPU (the product path has no CPU code).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from . import ops
from ._lib import GaeError
from .graph import coo_to_csr_torch


def block_bounds(n: int, world: int) -> List[int]:
    """Contiguous equal blocks (last one takes the remainder)."""
    per = (n + world - 1) // world
    return [min(i * per, n) for i in range(world + 1)]


def all_to_all_v(out: torch.Tensor, inp: torch.Tensor, out_splits: List[int], in_splits: List[int], group=None):
    """Row-wise all-to-all-v (splits count rows along dim 0).  NCCL executes this as one grouped
    ncclSend/ncclRecv per peer."""
    dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)


@dataclass
class HaloPlan:
    """Static exchange plan of one partitioned CSR."""
    rank: int
    world: int
    bounds: List[int]
    n_local: int
    n_halo: int
    halo_ids: torch.Tensor            # int64 [H] global ids of the remote sources, ascending
    recv_counts: List[int]            # rows received from each peer (halo rows owned by it)
    send_counts: List[int]            # rows sent to each peer
    send_idx: torch.Tensor            # int64 [S] LOCAL row indices to pack, grouped by peer
    rowptr: torch.Tensor              # local CSR over [local | halo] columns
    col: torch.Tensor
    plan: Optional[ops.HubPlan] = None
    n_edges: int = 0

    @property
    def halo_owner(self) -> torch.Tensor:
        o = torch.repeat_interleave(torch.arange(self.world, device=self.halo_ids.device),
                                    torch.tensor(self.recv_counts, device=self.halo_ids.device))
        return o.to(torch.int32)


def build_halo_plan(src: torch.Tensor, dst: torch.Tensor, n_global: int, rank: int, world: int,
                    group=None, seg_len: int = ops.DEFAULT_SEG_LEN) -> HaloPlan:
    """`src`, `dst`: global int64 endpoints of the edges whose dst this rank owns."""
    dev = src.device
    bounds = block_bounds(n_global, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    n_local = hi - lo
    if dst.numel() and (int(dst.min()) < lo or int(dst.max()) >= hi):
        raise GaeError("build_halo_plan: an edge's dst is not owned by this rank")
    remote = (src < lo) | (src >= hi)
    halo_ids = torch.unique(src[remote])                      # sorted ascending == grouped by owner
    n_halo = int(halo_ids.numel())
    col = torch.where(remote, n_local + torch.searchsorted(halo_ids, src), src - lo)
    rowptr, col32 = coo_to_csr_torch(col, dst - lo, n_local, n_cols=n_local + n_halo)
    b = torch.tensor(bounds, device=dev, dtype=torch.int64)
    cuts = torch.searchsorted(halo_ids, b)                    # halo range owned by each peer
    recv_counts = (cuts[1:] - cuts[:-1]).tolist()
    # tell every owner which of its rows we need
    rc = torch.tensor(recv_counts, dtype=torch.int64, device=dev)
    sc = torch.empty_like(rc)
    dist.all_to_all_single(sc, rc, group=group)
    send_counts = sc.tolist()
    owner_of = torch.repeat_interleave(torch.arange(world, device=dev), rc)
    req = halo_ids - b[owner_of]                              # owner-local row index
    send_idx = torch.empty(int(sc.sum()), dtype=torch.int64, device=dev)
    all_to_all_v(send_idx, req, send_counts, recv_counts, group)
    plan = None
    if rowptr.is_cuda:
        plan = ops.build_hub_plan(rowptr, seg_len)
        ops.order_segments_by_source(plan, rowptr, col32)
    return HaloPlan(rank, world, bounds, n_local, n_halo, halo_ids, recv_counts, send_counts, send_idx, rowptr,
                    col32, plan, int(src.numel()))


class PartitionedSpMM:
    """Y_local = (A X)[my rows] with one halo exchange.  `X_ext` is the [n_local + n_halo, d]
    feature buffer; callers write their rows into `X_ext[:n_local]`."""

    def __init__(self, hp: HaloPlan, d: int, exchange: str = "nccl", group=None,
                 pack_fn: Optional[Callable] = None, spmm_fn: Optional[Callable] = None):
        self.hp, self.d, self.group = hp, d, group
        dev = hp.rowptr.device
        self.X_ext = ops.alloc_rows(hp.n_local + hp.n_halo, d, dev) if dev.type == "cuda" else \
            torch.zeros(hp.n_local + hp.n_halo, d)
        self.Y = ops.alloc_rows(hp.n_local, d, dev) if dev.type == "cuda" else torch.zeros(hp.n_local, d)
        self.send_buf = torch.empty((max(int(hp.send_idx.numel()), 1), d), dtype=torch.float32, device=dev)
        self.ws = hp.plan.workspace(d, dev) if hp.plan is not None else None
        self.pack_fn = pack_fn or (lambda X, idx, out: ops.gather_rows(X, idx, out=out))
        self.spmm_fn = spmm_fn or (lambda rp, col, X, plan, out, ws: ops.spmm(rp, col, X, plan, out=out, partial_ws=ws))
        self.exchange = exchange
        self._peer_ptrs = None
        if exchange in ("p2p", "push"):
            self._setup_p2p()
        elif exchange != "nccl":
            raise GaeError(f"unknown exchange '{exchange}'")

    # ---- local / halo views ----------------------------------------------------------------
    @property
    def X_local(self) -> torch.Tensor:
        return self.X_ext[: self.hp.n_local]

    @property
    def X_halo(self) -> torch.Tensor:
        return self.X_ext[self.hp.n_local:]

    # ---- exchange ----------------------------------------------------------------------------
    def exchange_halo(self) -> None:
        hp = self.hp
        if self.exchange == "nccl":
            m = int(hp.send_idx.numel())
            if m:
                self.pack_fn(self.X_local, hp.send_idx, self.send_buf[:m])
            all_to_all_v(self.X_halo, self.send_buf[:m], hp.recv_counts, hp.send_counts, self.group)
        elif self.exchange == "push":
            self._push_p2p()
        else:
            self._pull_p2p()

    def _setup_p2p(self) -> None:
        """Map every peer's X_ext through CUDA IPC (one handle exchange per buffer)."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        hp = self.hp
        handle = (ctypes.c_uint8 * 64)()
        off = ctypes.c_int64(0)
        _lib.check(lib.gae_ipc_get_handle(ctypes.c_void_p(self.X_ext.data_ptr()), ctypes.byref(handle),
                                          ctypes.byref(off)), "gae_ipc_get_handle")
        mine = (bytes(handle), int(off.value), int(self.X_ext.stride(0)))
        everyone = [None] * hp.world
        dist.all_gather_object(everyone, mine, group=self.group)
        ptrs, self._opened = [], []
        for r, (h, o, ld) in enumerate(everyone):
            if ld != self.X_ext.stride(0):
                raise GaeError("p2p exchange needs the same feature row stride on every rank")
            if r == hp.rank:
                ptrs.append(self.X_ext.data_ptr())
                continue
            buf = (ctypes.c_uint8 * 64).from_buffer_copy(h)
            base = ctypes.c_void_p()
            _lib.check(lib.gae_ipc_open_handle(ctypes.byref(buf), ctypes.byref(base)), "gae_ipc_open_handle")
            self._opened.append(base.value)
            ptrs.append(base.value + o)
        dev = self.X_ext.device
        self._peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        b = torch.tensor(hp.bounds, dtype=torch.int64, device=dev)
        owner = hp.halo_owner
        self._pull_owner = owner
        self._pull_idx = hp.halo_ids - b[owner.to(torch.int64)]
        self._sync_flag = torch.zeros(1, dtype=torch.float32, device=dev)
        if self.exchange == "push":
            # where my rows land in each peer's [local | halo] buffer: peer q keeps the rows owned
            # by rank r at halo offset cuts_q[r]; every rank tells every owner that offset
            cuts = torch.zeros(hp.world + 1, dtype=torch.int64, device=dev)
            cuts[1:] = torch.cumsum(torch.tensor(hp.recv_counts, dtype=torch.int64, device=dev), 0)
            mine_at_peer = torch.empty(hp.world, dtype=torch.int64, device=dev)
            dist.all_to_all_single(mine_at_peer, (cuts[:-1] + hp.n_local).contiguous(), group=self.group)
            sc = torch.tensor(hp.send_counts, dtype=torch.int64, device=dev)
            peer = torch.repeat_interleave(torch.arange(hp.world, device=dev), sc)
            starts = torch.zeros(hp.world, dtype=torch.int64, device=dev)
            starts[1:] = torch.cumsum(sc, 0)[:-1]
            within = torch.arange(int(sc.sum()), device=dev, dtype=torch.int64) - starts[peer]
            # Interleave the destinations (row k of every peer list, then row k+1, ...): with the
            # lists merely concatenated, all ranks push to peer 0 first, then peer 1, ... and one
            # GPU's NVLink ingress throttles the whole box (measured at 8 GPUs: 170 GB/s/rank).
            order = torch.argsort(within * hp.world + peer)
            self._push_peer = peer[order].to(torch.int32).contiguous()
            self._push_row = (mine_at_peer[peer] + within)[order].contiguous()
            self._push_src = hp.send_idx[order].contiguous()

    def _device_barrier(self) -> None:
        """Stream-ordered barrier: a 1-element all-reduce.  Unlike dist.barrier() it does not block
        the host, so the pull and the SpMM behind it are enqueued back to back."""
        dist.all_reduce(self._sync_flag, group=self.group)

    def _push_p2p(self) -> None:
        import ctypes
        from . import _lib
        hp = self.hp
        m = int(hp.send_idx.numel())
        # peers must be done reading their halo rows of the previous SpMM before we overwrite them
        self._device_barrier()
        if m:
            rc = _lib.load().gae_push_rows_p2p_f32(ctypes.c_void_p(self.X_ext.data_ptr()), self.X_ext.stride(0),
                                                   ctypes.c_void_p(self._push_src.data_ptr()),
                                                   ctypes.c_void_p(self._push_peer.data_ptr()),
                                                   ctypes.c_void_p(self._push_row.data_ptr()),
                                                   ctypes.c_void_p(self._peer_ptrs.data_ptr()), m,
                                                   self.X_ext.stride(0), self.d,
                                                   torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "gae_push_rows_p2p_f32")
        # every push has completed (kernel boundary) on every rank before anyone consumes its halo
        self._device_barrier()

    def _pull_p2p(self) -> None:
        import ctypes
        from . import _lib
        hp = self.hp
        # peers must have finished writing their X_local before we read it, and we must not
        # overwrite ours while peers still read: two barriers bracket the pull
        self._device_barrier()
        if hp.n_halo:
            halo = self.X_halo
            rc = _lib.load().gae_pull_rows_p2p_f32(ctypes.c_void_p(self._peer_ptrs.data_ptr()),
                                                   ctypes.c_void_p(self._pull_owner.data_ptr()),
                                                   ctypes.c_void_p(self._pull_idx.data_ptr()), hp.n_halo,
                                                   self.X_ext.stride(0), self.d, ctypes.c_void_p(halo.data_ptr()),
                                                   halo.stride(0), torch.cuda.current_stream().cuda_stream)
            _lib.check(rc, "gae_pull_rows_p2p_f32")
        self._device_barrier()

    # ---- the op --------------------------------------------------------------------------------
    def __call__(self) -> torch.Tensor:
        self.exchange_halo()
        hp = self.hp
        self.spmm_fn(hp.rowptr, hp.col, self.X_ext, hp.plan, self.Y, self.ws)
        return self.Y


# ------------------------------------------------------------------------------------------------
# distributed R-MAT workload (bench.py --gpus N)
# ------------------------------------------------------------------------------------------------

def route_edges(src: torch.Tensor, dst: torch.Tensor, key: torch.Tensor, bounds: List[int], group=None):
    """Send every edge to the rank owning `key` (its row).  Returns the (src, dst) this rank owns."""
    world = len(bounds) - 1
    dev = src.device
    b = torch.tensor(bounds[1:], dtype=torch.int64, device=dev)
    owner = torch.searchsorted(b, key, right=True)
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)
    rc = torch.empty_like(counts)
    dist.all_to_all_single(rc, counts, group=group)
    payload = torch.stack([src[order], dst[order]], dim=1).contiguous()
    out = torch.empty((int(rc.sum()), 2), dtype=torch.int64, device=dev)
    all_to_all_v(out, payload, rc.tolist(), counts.tolist(), group)
    return out[:, 0].contiguous(), out[:, 1].contiguous()


@dataclass
class RmatPartition:
    fwd_op: PartitionedSpMM
    bwd_op: PartitionedSpMM
    local_edges: int
    local_rows: int
    halo_rows: int
    exchange_desc: str
    total_edges: int = 0
    d: int = 64
    _pinned: dict = field(default_factory=dict)

    def fwd(self):
        return self.fwd_op()

    def bwd(self):
        return self.bwd_op()

    def e2e(self, steps: int):
        """Host-resident features: H2D of this rank's X / dY rows, the partitioned step, D2H of Y / dX."""
        dev = self.fwd_op.X_ext.device
        n, d = self.local_rows, self.d
        hx = torch.empty((n, d), dtype=torch.float32, pin_memory=True).copy_(self.fwd_op.X_local)
        hdy = torch.empty((n, d), dtype=torch.float32, pin_memory=True).copy_(self.bwd_op.X_local)
        hy = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
        hdx = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
        st = torch.cuda.current_stream()

        def step():
            self.fwd_op.X_local.copy_(hx, non_blocking=True)
            hy.copy_(self.fwd_op(), non_blocking=True)
            self.bwd_op.X_local.copy_(hdy, non_blocking=True)
            hdx.copy_(self.bwd_op(), non_blocking=True)

        step()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            step()
        e1.record(st)
        e1.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms)
        world = dist.get_world_size()
        return {"value": self.total_edges / (ms * 1e-3), "unit": "edges/s", "ms_per_step": ms, "steps": steps,
                "h2d_bytes_per_step": int(2 * n * d * 4 * world), "d2h_bytes_per_step": int(2 * n * d * 4 * world),
                "api": "PartitionedSpMM with pinned host X/dY in, Y/dX out per rank; graph + halo plan resident"}


def build_rmat_partition(scale: int, total_edges: int, seed: int, d: int, device, exchange: str = "auto",
                         group=None, stages: int = 1) -> RmatPartition:
    from . import synthetic
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = 1 << scale
    bounds = block_bounds(n, world)
    per = (total_edges + world - 1) // world
    first = rank * per
    count = max(0, min(per, total_edges - first))
    src, dst = synthetic.rmat_edges(scale, count, seed=seed, device=device, first_edge=first)
    requested = exchange
    if exchange == "auto":
        exchange = "push"         # one-sided push (posted NVLink stores) is the fastest mechanism (profiles/)
    # forward: rows = dst
    fs, fd = route_edges(src, dst, dst, bounds, group)
    hp_f = build_halo_plan(fs, fd, n, rank, world, group)
    del fs, fd
    # backward: rows = src (CSR of A^T), entries = dst
    bs, bd = route_edges(dst, src, src, bounds, group)        # payload (entry, row) = (dst, src)
    del src, dst
    hp_b = build_halo_plan(bs, bd, n, rank, world, group)
    del bs, bd
    torch.cuda.empty_cache()
    def make_ops(mode):
        return PartitionedSpMM(hp_f, d, mode, group), PartitionedSpMM(hp_b, d, mode, group)

    if requested == "auto":
        # CUDA IPC needs peer access between every pair of GPUs; agree collectively, else use NCCL
        try:
            fwd_op, bwd_op = make_ops("push")
            ok = 1
        except Exception as exc:  # noqa: BLE001
            fwd_op = bwd_op = None
            ok = 0
            if rank == 0:
                print(f"[gae_dgl_b200.parallel] p2p exchange unavailable ({exc}); using NCCL all-to-all-v")
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) == 0:
            exchange = "nccl"
            del fwd_op, bwd_op
            fwd_op, bwd_op = make_ops("nccl")
    else:
        fwd_op, bwd_op = make_ops(exchange)
    lo = bounds[rank]
    fwd_op.X_local.copy_(synthetic.hashed_normal(hp_f.n_local, d, 2, device=device, first_row=lo))
    bwd_op.X_local.copy_(synthetic.hashed_normal(hp_b.n_local, d, 3, device=device, first_row=lo))
    if stages > 1 and exchange in ("push", "nccl"):
        # experimental (parallel_staged.py): B exchange stages overlapped with B row-block SpMMs
        from .parallel_staged import StagedPartitionedSpMM
        fwd_op = _StagedAdapter(StagedPartitionedSpMM(fwd_op, stages, group))
        bwd_op = _StagedAdapter(StagedPartitionedSpMM(bwd_op, stages, group))
    desc = {"nccl": "pack + NCCL all-to-all-v of deduplicated halo rows, per SpMM",
            "push": "one-sided push of deduplicated halo rows into peer HBM (CUDA IPC, posted NVLink stores), per SpMM",
            "p2p": "one-sided pull of deduplicated halo rows from peer HBM (CUDA IPC over NVLink), per SpMM"}[exchange]
    if stages > 1:
        desc += f"; {stages} exchange stages pipelined with row-block SpMMs"
    return RmatPartition(fwd_op, bwd_op, hp_f.n_edges, hp_f.n_local, hp_f.n_halo, desc, total_edges, d)


class _StagedAdapter:
    """Gives a StagedPartitionedSpMM the attribute surface RmatPartition uses (X_ext, X_local, call)."""

    def __init__(self, staged):
        self.staged = staged
        self.X_ext = staged.base.X_ext

    @property
    def X_local(self):
        return self.staged.base.X_local

    def __call__(self):
        return self.staged()
