"""Staged halo exchange pipelined with row-block SpMM (multi-GPU).

NOT on the default path yet: `parallel.PartitionedSpMM` (exchange everything, then one SpMM) is what
`bench.py --gpus N` runs and what was measured in round 1.  At 8 GPUs that serialises a 2.5 ms
exchange (1.76 GB of halo rows per rank at NVLink ingress speed -- the volume is already
deduplicated, so it cannot shrink) with a 3.5 ms local SpMM.  This module overlaps them:

  * the rank's destination rows are cut into B contiguous blocks of equal edge count;
  * every halo row is tagged with the FIRST block that reads it; stage s of the exchange delivers
    exactly the rows tagged s (each row still crosses NVLink once; the halo region of the
    [local | halo] buffer keeps its layout, so the local CSR needs no re-indexing);
  * block b's SpMM starts as soon as stage b has landed, while stage b+1 is in flight on a side
    stream: T ~ stage_0 + sum_b spmm_b instead of sum_s stage_s + sum_b spmm_b.  R-MAT halos are
    front-loaded (popular sources are hit by the first block already), so B = 4..8 is the useful
    range; results are bit-identical to the unstaged op for hub-free blocks and equal up to the
    hub-segment boundaries otherwise (each block has its own plan).

The planning logic (block cut, first-use tags, per-stage send / receive lists) runs on any device
and is covered by the gloo CPU tests with test doubles for the kernels, including the invariant that
block b only reads halo rows delivered by stages <= b (the halo is poisoned with NaN beforehand).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from . import ops
from ._lib import GaeError
from .parallel import HaloPlan, PartitionedSpMM, all_to_all_v


@dataclass
class StagePlan:
    n_stages: int
    row_bounds: List[int]                 # local row blocks, balanced by edge count; len = n_stages + 1
    sub_rowptr: List[torch.Tensor]        # per block: rowptr rebased to 0
    sub_col: List[torch.Tensor]           # per block: view into the local CSR's col array
    sub_plan: List[Optional[ops.HubPlan]]
    halo_stage: torch.Tensor              # int64 [n_halo]: first block that reads each halo row
    recv_pos: List[torch.Tensor]          # per stage: halo positions delivered, ascending (= grouped by owner)
    recv_counts: List[List[int]]          # per stage, per peer
    send_idx: List[torch.Tensor]          # per stage: LOCAL rows to send, grouped by peer
    send_counts: List[List[int]]
    send_stage: torch.Tensor              # int64 [S]: stage of every entry of HaloPlan.send_idx


def edge_balanced_bounds(rowptr: torch.Tensor, n_blocks: int) -> List[int]:
    """Contiguous row blocks holding ~equal numbers of edges (the last block takes the remainder)."""
    n = rowptr.numel() - 1
    total = int(rowptr[-1])
    if n_blocks <= 1 or n == 0:
        return [0, n]
    targets = torch.tensor([total * k // n_blocks for k in range(1, n_blocks)], dtype=rowptr.dtype, device=rowptr.device)
    cuts = torch.searchsorted(rowptr, targets, right=False).clamp_(0, n).tolist()
    bounds = [0]
    for c in cuts:
        bounds.append(max(int(c), bounds[-1]))
    bounds.append(n)
    return bounds


def build_stage_plan(hp: HaloPlan, n_stages: int, group=None, seg_len: int = ops.DEFAULT_SEG_LEN) -> StagePlan:
    dev = hp.rowptr.device
    n_local, n_halo, world = hp.n_local, hp.n_halo, hp.world
    bounds = edge_balanced_bounds(hp.rowptr, n_stages)
    n_stages = len(bounds) - 1
    b_t = torch.tensor(bounds, dtype=torch.int64, device=dev)
    # first block that reads each halo row
    deg = hp.rowptr[1:] - hp.rowptr[:-1]
    edge_row = torch.repeat_interleave(torch.arange(n_local, device=dev, dtype=torch.int64), deg)
    edge_block = torch.bucketize(edge_row, b_t[1:], right=True)
    col = hp.col.to(torch.int64)
    remote = col >= n_local
    halo_stage = torch.full((n_halo,), n_stages, dtype=torch.int64, device=dev)
    if n_halo:
        halo_stage.scatter_reduce_(0, col[remote] - n_local, edge_block[remote], reduce="amin", include_self=True)
        if int(halo_stage.max()) >= n_stages:
            raise GaeError("build_stage_plan: a halo row is referenced by no edge")
    # what I receive per stage (ascending halo position == grouped by owner, owners being contiguous id ranges)
    cuts = torch.zeros(world + 1, dtype=torch.int64, device=dev)
    cuts[1:] = torch.cumsum(torch.tensor(hp.recv_counts, dtype=torch.int64, device=dev), 0)
    recv_pos, recv_counts = [], []
    for s in range(n_stages):
        pos = torch.nonzero(halo_stage == s, as_tuple=False).flatten()
        recv_pos.append(pos)
        edges = torch.searchsorted(pos, cuts)
        recv_counts.append((edges[1:] - edges[:-1]).tolist())
    # tell every owner the stage of each row it sends me (same order as the original request lists)
    send_stage = torch.empty(int(hp.send_idx.numel()), dtype=torch.int64, device=dev)
    all_to_all_v(send_stage, halo_stage, hp.send_counts, hp.recv_counts, group)
    peer_of_send = torch.repeat_interleave(torch.arange(world, device=dev),
                                           torch.tensor(hp.send_counts, dtype=torch.int64, device=dev))
    send_idx, send_counts = [], []
    for s in range(n_stages):
        m = send_stage == s
        send_idx.append(hp.send_idx[m].contiguous())
        send_counts.append(torch.bincount(peer_of_send[m], minlength=world).tolist())
    # row-block sub-CSRs (views of the local CSR; only rowptr is rebased)
    sub_rowptr, sub_col, sub_plan = [], [], []
    for s in range(n_stages):
        r0, r1 = bounds[s], bounds[s + 1]
        e0, e1 = int(hp.rowptr[r0]), int(hp.rowptr[r1])
        rp = (hp.rowptr[r0:r1 + 1] - e0).contiguous()
        cl = hp.col[e0:e1]
        plan = None
        if rp.is_cuda and r1 > r0:
            plan = ops.build_hub_plan(rp, seg_len)
            ops.order_segments_by_source(plan, rp, cl)
        sub_rowptr.append(rp)
        sub_col.append(cl)
        sub_plan.append(plan)
    return StagePlan(n_stages, bounds, sub_rowptr, sub_col, sub_plan, halo_stage, recv_pos, recv_counts, send_idx,
                     send_counts, send_stage)


class StagedPartitionedSpMM:
    """Y_local = (A X)[my rows]: B exchange stages overlapped with B row-block SpMMs.

    Wraps a `PartitionedSpMM` (its [local | halo] buffer, output, peer mappings and interleaved push
    lists are reused).  `exchange` is the wrapped op's: "push" runs one push kernel per stage over the
    matching slice of the interleaved lists; "nccl" packs, all-to-all-v's and scatters per stage."""

    def __init__(self, base: PartitionedSpMM, n_stages: int, group=None, overlap: Optional[bool] = None):
        if base.exchange not in ("push", "nccl"):
            raise GaeError("staged exchange supports the 'push' and 'nccl' mechanisms")
        self.base, self.group = base, group
        self.sp = build_stage_plan(base.hp, n_stages, group)
        dev = base.X_ext.device
        self.overlap = (dev.type == "cuda") if overlap is None else overlap
        self.ws = [p.workspace(base.d, dev) if p is not None else None for p in self.sp.sub_plan]
        self._recv_buf = [torch.empty((max(int(p.numel()), 1), base.d), dtype=torch.float32, device=dev)
                          for p in self.sp.recv_pos] if base.exchange == "nccl" else None
        self._send_buf = [torch.empty((max(int(i.numel()), 1), base.d), dtype=torch.float32, device=dev)
                          for i in self.sp.send_idx] if base.exchange == "nccl" else None
        if base.exchange == "push":
            # slices of the wrapped op's interleaved push lists, stage by stage (order preserved)
            order_stage = self._push_stage_of_entries()
            self._push = []
            for s in range(self.sp.n_stages):
                m = order_stage == s
                self._push.append((base._push_src[m].contiguous(), base._push_peer[m].contiguous(),
                                   base._push_row[m].contiguous()))
        self._comm = torch.cuda.Stream(device=dev) if self.overlap else None

    def _push_stage_of_entries(self) -> torch.Tensor:
        """Stage of every entry of base._push_* (they are hp.send_idx permuted by the interleaving)."""
        b, hp = self.base, self.base.hp
        dev = b.X_ext.device
        sc = torch.tensor(hp.send_counts, dtype=torch.int64, device=dev)
        peer = torch.repeat_interleave(torch.arange(hp.world, device=dev), sc)
        starts = torch.zeros(hp.world, dtype=torch.int64, device=dev)
        starts[1:] = torch.cumsum(sc, 0)[:-1]
        within = torch.arange(int(sc.sum()), device=dev, dtype=torch.int64) - starts[peer]
        order = torch.argsort(within * hp.world + peer)          # the same permutation _setup_p2p applies
        return self.sp.send_stage[order]

    # ---- one exchange stage -----------------------------------------------------------------
    def exchange_stage(self, s: int) -> None:
        b, sp = self.base, self.sp
        if b.exchange == "nccl":
            m = int(sp.send_idx[s].numel())
            k = int(sp.recv_pos[s].numel())
            if m:
                b.pack_fn(b.X_local, sp.send_idx[s], self._send_buf[s][:m])
            all_to_all_v(self._recv_buf[s][:k], self._send_buf[s][:m], sp.recv_counts[s], sp.send_counts[s], self.group)
            if k:
                b.X_halo.index_copy_(0, sp.recv_pos[s], self._recv_buf[s][:k])
            return
        import ctypes
        from . import _lib
        src, peer, row = self._push[s]
        m = int(src.numel())
        if m:
            rc = _lib.load().gae_push_rows_p2p_f32(ctypes.c_void_p(b.X_ext.data_ptr()), b.X_ext.stride(0),
                                                   ctypes.c_void_p(src.data_ptr()), ctypes.c_void_p(peer.data_ptr()),
                                                   ctypes.c_void_p(row.data_ptr()),
                                                   ctypes.c_void_p(b._peer_ptrs.data_ptr()), m, b.X_ext.stride(0), b.d,
                                                   ops._stream())
            _lib.check(rc, "gae_push_rows_p2p_f32")
        b._device_barrier()          # every rank's stage-s rows have landed before anyone's block s reads them

    def spmm_block(self, s: int) -> None:
        b, sp = self.base, self.sp
        r0, r1 = sp.row_bounds[s], sp.row_bounds[s + 1]
        if r1 > r0:
            b.spmm_fn(sp.sub_rowptr[s], sp.sub_col[s], b.X_ext, sp.sub_plan[s], b.Y[r0:r1], self.ws[s])

    # ---- the op -------------------------------------------------------------------------------
    def __call__(self) -> torch.Tensor:
        b, sp = self.base, self.sp
        if not self.overlap:                       # sequential form (CPU tests; also a debugging aid)
            if b.exchange == "push":
                b._device_barrier()
            for s in range(sp.n_stages):
                self.exchange_stage(s)
                self.spmm_block(s)
            return b.Y
        cur = torch.cuda.current_stream()
        self._comm.wait_stream(cur)                # X_local is written, the previous SpMM has read its halo
        landed = []
        with torch.cuda.stream(self._comm):
            if b.exchange == "push":
                b._device_barrier()                # ... on every rank, before any halo row is overwritten
            for s in range(sp.n_stages):
                self.exchange_stage(s)
                ev = torch.cuda.Event()
                ev.record(self._comm)
                landed.append(ev)
        for s in range(sp.n_stages):
            cur.wait_event(landed[s])
            self.spmm_block(s)
        return b.Y
