"""1-D vertex partition of the encoder SpMM across the GPUs of one box.

Rank r owns the contiguous block of destination rows [r*V/P, (r+1)*V/P), their in-edge CSR and
the feature rows of those vertices.  Column ids are remapped to [local | halo]: halo = sorted
unique remote sources, which (blocks being contiguous) are already grouped by owner.  One
exchange per SpMM moves each needed remote row exactly once.  Mechanisms:

  halo : (default) `HaloSpMM` -- the whole operator behind ONE C call (gae_halo_spmm_f32,
         csrc/halo.cu): a persistent one-sided push kernel delivers the halo rows stage by stage
         into the peers' buffers (CUDA IPC, posted NVLink stores) and publishes device-side flags;
         the row-block SpMM of stage s starts when its flags have landed, so the transfer of the
         later stages overlaps the aggregation of the earlier ones.  No collective, no host sync.
  nccl : pack (gae_gather_rows_f32) -> all-to-all-v (NCCL grouped send/recv over NVLink)
         -> rows land directly in the halo region; then one SpMM.  Fallback when IPC is unavailable.
  push / p2p : the round-1 one-sided push / pull bracketed by two stream-ordered NCCL barriers,
         kept for comparison (bench.py --exchange push).

Planning (halo ids, [local | halo] column remap, first-use stage tags, interleaved send lists) is
done by the host helpers of the C ABI (gae_halo_*_host); torch.distributed only carries the
request lists between the ranks at set-up time.
The backward SpMM (dX = A^T dY) uses the same machinery on CSR(A^T), partitioned by source.
Graph generation is distributed as well: every rank draws 1/P of the R-MAT edge stream and
routes each edge to the owner of its row.

`pack_fn` / `spmm_fn` default to the CUDA kernels; the gloo CPU tests inject test doubles for
them so the planning / exchange logic runs without a GPU (the product path has no CPU code).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import GaeError, HaloBlockStruct, HaloExchangeStruct
from .graph import coo_to_csr_torch

DEFAULT_STAGES = 8


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


def _np_ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def halo_plan_host(src_global: np.ndarray, bounds: List[int], rank: int):
    """gae_halo_plan_count_host / _fill_host on host arrays: (halo_ids int64 [H] ascending,
    col_local int32 [E] in [local | halo], recv_counts [world])."""
    lib = _lib.load()
    world = len(bounds) - 1
    src_global = np.ascontiguousarray(src_global, dtype=np.int64)
    b = np.asarray(bounds, dtype=np.int64)
    n_halo = ctypes.c_int64(0)
    recv = np.zeros(world, dtype=np.int64)
    _lib.check(lib.gae_halo_plan_count_host(_np_ptr(src_global), src_global.size, _np_ptr(b), world, rank,
                                            ctypes.byref(n_halo), _np_ptr(recv)), "gae_halo_plan_count_host")
    halo_ids = np.zeros(max(n_halo.value, 1), dtype=np.int64)
    col_local = np.zeros(max(src_global.size, 1), dtype=np.int32)
    _lib.check(lib.gae_halo_plan_fill_host(_np_ptr(src_global), src_global.size, _np_ptr(b), world, rank,
                                           _np_ptr(halo_ids), _np_ptr(col_local)), "gae_halo_plan_fill_host")
    return halo_ids[:n_halo.value], col_local[:src_global.size], recv.tolist()


def build_halo_plan(src: torch.Tensor, dst: torch.Tensor, n_global: int, rank: int, world: int,
                    group=None, seg_len: int = ops.DEFAULT_SEG_LEN) -> HaloPlan:
    """`src`, `dst`: global int64 endpoints of the edges whose dst this rank owns."""
    dev = src.device
    bounds = block_bounds(n_global, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    n_local = hi - lo
    if dst.numel() and (int(dst.min()) < lo or int(dst.max()) >= hi):
        raise GaeError("build_halo_plan: an edge's dst is not owned by this rank")
    halo_np, col_np, recv_counts = halo_plan_host(src.cpu().numpy(), bounds, rank)
    halo_ids = torch.from_numpy(halo_np).to(dev)
    n_halo = int(halo_ids.numel())
    col = torch.from_numpy(col_np).to(dev).to(torch.int64)
    rowptr, col32 = coo_to_csr_torch(col, dst - lo, n_local, n_cols=n_local + n_halo)
    del col
    b = torch.tensor(bounds, device=dev, dtype=torch.int64)
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


# ------------------------------------------------------------------------------------------------
# staged exchange plan (first-use tags, row blocks, interleaved send lists)
# ------------------------------------------------------------------------------------------------

@dataclass
class StagePlan:
    """Stages of the exchange and the SpMM pieces that consume them.  kind "blocks": stage s = the halo rows
    first read by row block s, piece s = the rows of that block over all their edges.  kind "classes":
    stage s = the halo rows of popularity class s (the buffer keeps them in that order), piece s = ALL rows
    over the edges whose source is in class s (local sources count as class 0), added into Y for s > 0."""
    kind: str
    n_stages: int
    row_bounds: List[int]                 # kind "blocks": local row blocks; kind "classes": [0, n_local]
    piece_rows: List[tuple]               # per piece: (row0, n_rows)
    sub_rowptr: List[torch.Tensor]        # per piece: rowptr rebased to 0
    sub_col: List[torch.Tensor]           # per piece: columns in the [local | halo] buffer
    sub_plan: List[Optional[ops.HubPlan]]
    accumulate: List[bool]
    halo_stage: torch.Tensor              # int32 [n_halo] (host): stage of each halo row (HaloPlan.halo_ids order)
    halo_pos: torch.Tensor                # int64 [n_halo] (host): its row in my [local | halo] buffer
    send_stage: torch.Tensor              # int32 [S] (host): stage of every entry of HaloPlan.send_idx
    push_src: torch.Tensor                # int64 [S]: local rows to send, sorted by stage, peers interleaved
    push_peer: torch.Tensor               # int32 [S]
    push_dst: torch.Tensor                # int64 [S]: row in the destination's [local | halo] buffer
    stage_ptr: np.ndarray                 # int64 [n_stages + 1] (host)
    # kind "fold" only: the halo part of the buffer holds n_ext rows (copied hot / cold sources, then the
    # folded rows the owners sum for me); behind it n_pre staging rows for the sums I owe my peers
    n_ext: int = -1                       # halo rows in the buffer (-1: HaloPlan.n_halo)
    n_pre: int = 0
    pre_stage: int = 0
    pre_rowptr: Optional[torch.Tensor] = None   # int64 [n_pre + 1]
    pre_col: Optional[torch.Tensor] = None      # int32: my local rows
    pre_plan: Optional[ops.HubPlan] = None
    stats: Optional[dict] = None


def edge_balanced_bounds(rowptr: np.ndarray, n_blocks: int) -> List[int]:
    """Contiguous row blocks holding ~equal numbers of edges (the last block takes the remainder)."""
    n = rowptr.shape[0] - 1
    total = int(rowptr[-1])
    if n_blocks <= 1 or n == 0:
        return [0, n]
    targets = np.asarray([total * k // n_blocks for k in range(1, n_blocks)], dtype=np.int64)
    cuts = np.clip(np.searchsorted(rowptr, targets, side="left"), 0, n).tolist()
    bounds = [0]
    for c in cuts:
        bounds.append(max(int(c), bounds[-1]))
    bounds.append(n)
    return bounds


def _send_lists(hp: HaloPlan, halo_stage: np.ndarray, halo_pos: np.ndarray, n_stages: int, group):
    """Tell every owner the stage and the destination row of each row it sends me (same order as the request
    lists), then build my own staged, peer-interleaved send lists (gae_halo_push_lists_host)."""
    lib = _lib.load()
    dev = hp.rowptr.device
    m = int(hp.send_idx.numel())
    send_stage = torch.empty(m, dtype=torch.int32, device=dev)
    all_to_all_v(send_stage, torch.from_numpy(halo_stage).to(dev), hp.send_counts, hp.recv_counts, group)
    send_dst = torch.empty(m, dtype=torch.int64, device=dev)
    all_to_all_v(send_dst, torch.from_numpy(halo_pos).to(dev), hp.send_counts, hp.recv_counts, group)
    send_idx = np.ascontiguousarray(hp.send_idx.cpu().numpy(), dtype=np.int64)
    sst = np.ascontiguousarray(send_stage.cpu().numpy(), dtype=np.int32)
    sdst = np.ascontiguousarray(send_dst.cpu().numpy(), dtype=np.int64)
    sc = np.asarray(hp.send_counts, dtype=np.int64)
    o_src, o_peer, o_dst = (np.zeros(max(m, 1), dtype=np.int64), np.zeros(max(m, 1), dtype=np.int32),
                            np.zeros(max(m, 1), dtype=np.int64))
    stage_ptr = np.zeros(n_stages + 1, dtype=np.int64)
    _lib.check(lib.gae_halo_push_lists_host(_np_ptr(send_idx), _np_ptr(sst), _np_ptr(sdst), _np_ptr(sc), None, hp.world,
                                            n_stages, _np_ptr(o_src), _np_ptr(o_peer), _np_ptr(o_dst), _np_ptr(stage_ptr)),
               "gae_halo_push_lists_host")
    return (torch.from_numpy(sst.copy()), torch.from_numpy(o_src[:m]).to(dev), torch.from_numpy(o_peer[:m]).to(dev),
            torch.from_numpy(o_dst[:m]).to(dev), stage_ptr)


def _piece_plan(rowptr: torch.Tensor, col: torch.Tensor, seg_len: int) -> Optional[ops.HubPlan]:
    if not rowptr.is_cuda or rowptr.numel() <= 1:
        return None
    plan = ops.build_hub_plan(rowptr, seg_len)
    ops.order_segments_by_source(plan, rowptr, col)
    return plan


def build_stage_plan(hp: HaloPlan, n_stages: int, group=None, seg_len: int = ops.DEFAULT_SEG_LEN) -> StagePlan:
    """Row-block stages: cut the rank's rows into `n_stages` edge-balanced blocks and tag every halo row with
    the first block that reads it (gae_halo_stage_tags_host).  The halo keeps the HaloPlan's layout."""
    lib = _lib.load()
    n_stages = max(1, min(int(n_stages), 32))
    rp = hp.rowptr.cpu().numpy()
    cl = hp.col.cpu().numpy()
    bounds = edge_balanced_bounds(rp, n_stages)
    n_stages = len(bounds) - 1
    rb = np.asarray(bounds, dtype=np.int64)
    halo_stage = np.zeros(max(hp.n_halo, 1), dtype=np.int32)
    _lib.check(lib.gae_halo_stage_tags_host(_np_ptr(rp), _np_ptr(cl), hp.n_local, hp.n_halo, _np_ptr(rb), n_stages,
                                            _np_ptr(halo_stage)), "gae_halo_stage_tags_host")
    halo_stage = halo_stage[:hp.n_halo]
    halo_pos = hp.n_local + np.arange(hp.n_halo, dtype=np.int64)
    sst, p_src, p_peer, p_dst, stage_ptr = _send_lists(hp, halo_stage, halo_pos, n_stages, group)
    # row-block sub-CSRs (views of the local CSR; only rowptr is rebased)
    sub_rowptr, sub_col, sub_plan = [], [], []
    for s in range(n_stages):
        r0, r1 = bounds[s], bounds[s + 1]
        e0, e1 = int(rp[r0]), int(rp[r1])
        srp = (hp.rowptr[r0:r1 + 1] - e0).contiguous()
        scl = hp.col[e0:e1]
        sub_rowptr.append(srp)
        sub_col.append(scl)
        sub_plan.append(_piece_plan(srp, scl, seg_len) if r1 > r0 else None)
    return StagePlan("blocks", n_stages, bounds, [(bounds[s], bounds[s + 1] - bounds[s]) for s in range(n_stages)],
                     sub_rowptr, sub_col, sub_plan, [False] * n_stages, torch.from_numpy(halo_stage.copy()),
                     torch.from_numpy(halo_pos), sst, p_src, p_peer, p_dst, stage_ptr)


DEFAULT_HOT_THRESHOLDS = (8,)


def build_class_plan(hp: HaloPlan, thresholds=DEFAULT_HOT_THRESHOLDS, group=None,
                     seg_len: int = ops.DEFAULT_SEG_LEN) -> StagePlan:
    """Popularity-class stages.  A halo row referenced by >= thresholds[0] of this rank's edges is class 0,
    by >= thresholds[1] class 1, ..., the rest the last class (thresholds descending).  On skewed graphs a
    small class 0 carries most of the edges (R-MAT, 8 ranks: the 21 % of the halo rows referenced >= 8 times
    carry 86 % of a rank's edges), so it is delivered first and the bulk of the aggregation -- local and
    class-0 sources, piece 0 -- runs while the long tail of rarely used rows is still in flight; the tail
    classes are then added into Y (pieces s > 0, accumulate).  The halo part of the buffer is laid out class
    by class (ascending global id inside a class) and every piece is its own CSR over all local rows."""
    dev = hp.rowptr.device
    n_local, n_halo = hp.n_local, hp.n_halo
    thresholds = sorted((int(t) for t in thresholds), reverse=True)
    K = len(thresholds) + 1
    col = hp.col.to(torch.int64)
    is_halo = col >= n_local
    cnt = torch.bincount(col[is_halo] - n_local, minlength=n_halo)
    cls = torch.full((n_halo,), K - 1, dtype=torch.int64, device=dev)
    for t in thresholds:
        cls -= (cnt >= t).to(torch.int64)
    order = torch.argsort(cls, stable=True)                    # class-major, ascending id inside a class
    newpos = torch.empty(n_halo, dtype=torch.int64, device=dev)
    newpos[order] = torch.arange(n_halo, dtype=torch.int64, device=dev)
    deg = hp.rowptr[1:] - hp.rowptr[:-1]
    rows = torch.repeat_interleave(torch.arange(n_local, dtype=torch.int64, device=dev), deg)
    h = (col - n_local).clamp_(min=0)
    ecls = torch.where(is_halo, cls[h] if n_halo else torch.zeros_like(col), torch.zeros_like(col))
    newcol = torch.where(is_halo, n_local + (newpos[h] if n_halo else torch.zeros_like(col)), col)
    del col, h, deg
    sub_rowptr, sub_col, sub_plan = [], [], []
    for k in range(K):
        m = ecls == k
        srp, scl = coo_to_csr_torch(newcol[m], rows[m], n_local, n_cols=n_local + n_halo)
        sub_rowptr.append(srp)
        sub_col.append(scl)
        sub_plan.append(_piece_plan(srp, scl, seg_len))
    del rows, newcol, ecls, is_halo
    halo_stage = cls.to(torch.int32).cpu().numpy()
    halo_pos = (n_local + newpos).cpu().numpy()
    sst, p_src, p_peer, p_dst, stage_ptr = _send_lists(hp, halo_stage, halo_pos, K, group)
    return StagePlan("classes", K, [0, n_local], [(0, n_local)] * K, sub_rowptr, sub_col, sub_plan,
                     [k > 0 for k in range(K)], torch.from_numpy(halo_stage.copy()), torch.from_numpy(halo_pos.copy()),
                     sst, p_src, p_peer, p_dst, stage_ptr)


DEFAULT_FOLD = (16, 2)       # (hot, fold): see build_fold_plan


def build_fold_plan(hp: HaloPlan, hot: int = DEFAULT_FOLD[0], fold: int = DEFAULT_FOLD[1], group=None,
                    seg_len: int = ops.DEFAULT_SEG_LEN) -> StagePlan:
    """Popularity classes + sender-side FOLDING of the long tail (kind "fold", the default).

    A remote source referenced by >= `hot` of my edges is copied as before (stage 0: few rows, most of the
    edges).  The other remote edges mostly run from rarely used sources into my HUB rows, so for every (owner q,
    my row j) with >= `fold` such edges I do not fetch the sources: q sums them for me over its own rows and
    sends ONE folded row, which my CSR references as a single column.  What is left (cold sources of sparse
    rows) is copied.  A vertex cover of the remote bipartite graph, chosen by two thresholds; R-MAT at 8 ranks:
    0.56 of the deduplicated halo volume (hot 16, fold 2), and the folded edges leave my aggregation for the
    owner's.  Stage 0 = hot rows, stage 1 = cold rows + folded rows; piece 0 = local + hot sources, piece 1
    (accumulate) = cold sources + folded rows.  The sums I owe my peers are one more SpMM over my local rows
    (pre_rowptr / pre_col) into staging rows behind my halo region; the push kernel reads them from there."""
    lib = _lib.load()
    dev = hp.rowptr.device
    world, n_local, n_halo = hp.world, hp.n_local, hp.n_halo
    hot, fold = max(int(hot), 1), max(int(fold), 2)
    i64 = dict(dtype=torch.int64, device=dev)
    nh = max(n_halo, 1)
    col = hp.col.to(torch.int64)
    is_halo = col >= n_local
    h = (col - n_local).clamp_(min=0)
    cnt = torch.bincount(h[is_halo], minlength=nh)
    hot_row = cnt >= hot
    if n_halo == 0:
        hot_row[:] = False
    owner = torch.zeros(nh, **i64)
    owner[:n_halo] = torch.repeat_interleave(torch.arange(world, **i64), torch.tensor(hp.recv_counts, **i64))
    b = torch.tensor(hp.bounds, **i64)
    deg = hp.rowptr[1:] - hp.rowptr[:-1]
    rows = torch.repeat_interleave(torch.arange(n_local, **i64), deg)
    del deg, cnt
    e_hot = is_halo & hot_row[h]
    res = is_halo & ~e_hot
    res_rows, res_h = rows[res], h[res]
    del res
    key = owner[res_h] * n_local + res_rows
    uk, inv, kc = torch.unique(key, return_inverse=True, return_counts=True)
    del key
    folded_key = kc >= fold
    e_fold = folded_key[inv] if inv.numel() else torch.zeros(0, dtype=torch.bool, device=dev)
    cold_row = torch.zeros(nh, dtype=torch.bool, device=dev)
    cold_row[res_h[~e_fold]] = True
    n_hot, n_cold, n_fold = int(hot_row.sum()), int(cold_row.sum()), int(folded_key.sum())
    n_ext = n_hot + n_cold + n_fold
    pos = torch.full((nh,), -1, **i64)
    pos[hot_row] = torch.arange(n_hot, **i64)
    pos[cold_row] = n_hot + torch.arange(n_cold, **i64)
    # my two pieces
    m0 = ~is_halo | e_hot
    c0 = torch.where(is_halo, n_local + pos[h], col)[m0]
    rp0, cl0 = coo_to_csr_torch(c0, rows[m0], n_local, n_cols=n_local + n_ext)
    del m0, c0, rows, col, is_halo, e_hot, h
    fk = uk[folded_key]
    f_cols = n_local + n_hot + n_cold + torch.arange(n_fold, **i64)
    r1 = torch.cat([res_rows[~e_fold], fk % n_local])
    c1 = torch.cat([n_local + pos[res_h[~e_fold]], f_cols])
    rp1, cl1 = coo_to_csr_torch(c1, r1, n_local, n_cols=n_local + n_ext)
    n_cold_edges = int((~e_fold).sum())
    del r1, c1
    # requests: (a) rows to copy, in halo order = grouped by owner; (b) rows to fold, ascending key = grouped by owner
    sel = (hot_row | cold_row)[:n_halo]
    sel_ids = torch.nonzero(sel).flatten()
    copy_owner = owner[sel_ids]
    copy_idx = hp.halo_ids[sel_ids] - b[copy_owner]
    copy_stage = (~hot_row[sel_ids]).to(torch.int32)
    copy_dst = n_local + pos[sel_ids]
    fold_owner = torch.div(fk, max(n_local, 1), rounding_mode="floor")
    fold_len = kc[folded_key]
    slot_of_key = torch.cumsum(folded_key.to(torch.int64), 0) - 1
    e_slot = slot_of_key[inv[e_fold]] if inv.numel() else torch.zeros(0, **i64)
    e_order = torch.argsort(e_slot, stable=True)
    e_h = res_h[e_fold][e_order]
    fold_cols = (hp.halo_ids[e_h] - b[owner[e_h]]) if n_halo else torch.zeros(0, **i64)
    n_fold_edges = int(e_h.numel())
    del e_slot, e_order, e_h, inv, res_h, res_rows, e_fold
    cnts = torch.stack([torch.bincount(copy_owner, minlength=world), torch.bincount(fold_owner, minlength=world),
                        torch.zeros(world, **i64).index_add_(0, fold_owner, fold_len)], dim=1).contiguous()
    got = torch.empty_like(cnts)
    dist.all_to_all_single(got, cnts, group=group)
    mine, theirs = cnts.cpu().tolist(), got.cpu().tolist()

    def swap(t, k):
        out = torch.empty(sum(x[k] for x in theirs), dtype=t.dtype, device=dev)
        all_to_all_v(out, t.contiguous(), [x[k] for x in theirs], [x[k] for x in mine], group)
        return out

    g_idx, g_stage, g_dst = swap(copy_idx, 0), swap(copy_stage, 0), swap(copy_dst, 0)
    g_flen, g_fdst, g_fcols = swap(fold_len, 1), swap(f_cols, 1), swap(fold_cols, 2)
    # the sums I owe: one staging row per requested folded row, in peer order
    n_pre = int(g_flen.numel())
    pre_row0 = n_local + n_ext
    pre_rowptr = torch.zeros(n_pre + 1, **i64)
    pre_rowptr[1:] = torch.cumsum(g_flen, 0)
    pre_col = g_fcols.to(torch.int32)
    if n_pre and (int(g_fcols.min()) < 0 or int(g_fcols.max()) >= n_local):
        raise GaeError("build_fold_plan: a peer asked me to sum rows I do not own")
    # combined send lists per peer: [rows to copy | folded rows], then staged and interleaved
    c_first = np.concatenate([[0], np.cumsum([x[0] for x in theirs])]).astype(np.int64)
    f_first = np.concatenate([[0], np.cumsum([x[1] for x in theirs])]).astype(np.int64)
    idx_np, st_np, dst_np = g_idx.cpu().numpy(), g_stage.cpu().numpy(), g_dst.cpu().numpy()
    fdst_np = g_fdst.cpu().numpy()
    s_idx, s_stage, s_dst, s_counts = [], [], [], []
    for q in range(world):
        c0_, c1_, f0_, f1_ = c_first[q], c_first[q + 1], f_first[q], f_first[q + 1]
        s_idx += [idx_np[c0_:c1_], pre_row0 + np.arange(f0_, f1_, dtype=np.int64)]
        s_stage += [st_np[c0_:c1_], np.ones(f1_ - f0_, dtype=np.int32)]
        s_dst += [dst_np[c0_:c1_], fdst_np[f0_:f1_]]
        s_counts.append(int(c1_ - c0_ + f1_ - f0_))
    send_idx = np.ascontiguousarray(np.concatenate(s_idx), dtype=np.int64)
    sst = np.ascontiguousarray(np.concatenate(s_stage), dtype=np.int32)
    sdst = np.ascontiguousarray(np.concatenate(s_dst), dtype=np.int64)
    sc = np.asarray(s_counts, dtype=np.int64)
    m = int(send_idx.size)
    o_src, o_peer, o_dst = (np.zeros(max(m, 1), dtype=np.int64), np.zeros(max(m, 1), dtype=np.int32),
                            np.zeros(max(m, 1), dtype=np.int64))
    stage_ptr = np.zeros(3, dtype=np.int64)
    _lib.check(lib.gae_halo_push_lists_host(_np_ptr(send_idx), _np_ptr(sst), _np_ptr(sdst), _np_ptr(sc), None, world, 2,
                                            _np_ptr(o_src), _np_ptr(o_peer), _np_ptr(o_dst), _np_ptr(stage_ptr)),
               "gae_halo_push_lists_host")
    halo_stage = torch.where(hot_row, 0, torch.where(cold_row, 1, -1))[:n_halo].to(torch.int32).cpu()
    halo_pos = torch.where(pos >= 0, n_local + pos, -1)[:n_halo].cpu()
    stats = {"halo_rows_dedup": n_halo, "hot_rows": n_hot, "cold_rows": n_cold, "folded_rows_in": n_fold,
             "rows_in": n_ext, "rows_out": m, "folded_rows_out": n_pre, "folded_edges_out": int(pre_col.numel()),
             "folded_edges_in": n_fold_edges, "cold_edges": n_cold_edges, "hot": hot, "fold": fold}
    return StagePlan("fold", 2, [0, n_local], [(0, n_local)] * 2, [rp0, rp1], [cl0, cl1],
                     [_piece_plan(rp0, cl0, seg_len), _piece_plan(rp1, cl1, seg_len)], [False, True], halo_stage, halo_pos,
                     torch.from_numpy(sst.copy()), torch.from_numpy(o_src[:m]).to(dev), torch.from_numpy(o_peer[:m]).to(dev),
                     torch.from_numpy(o_dst[:m]).to(dev), stage_ptr, n_ext=n_ext, n_pre=n_pre, pre_stage=1,
                     pre_rowptr=pre_rowptr, pre_col=pre_col,
                     pre_plan=_piece_plan(pre_rowptr, pre_col, seg_len) if n_pre else None, stats=stats)


# ------------------------------------------------------------------------------------------------
# peer memory (CUDA IPC) shared by the operators of one process
# ------------------------------------------------------------------------------------------------

FLAG_WORDS = 1024          # GAE_HALO_FLAG_WORDS
FLAG_BLOCKS = 64


class PeerMemory:
    """CUDA IPC mappings of this process: every peer allocation is opened once (opening the same
    handle twice in one process fails) and one flag pool serves all partitioned operators."""
    _by_group: Dict[int, "PeerMemory"] = {}

    @classmethod
    def get(cls, device, group=None) -> "PeerMemory":
        key = id(group)
        pm = cls._by_group.get(key)
        if pm is None or pm.device != device:
            pm = cls(device, group)
            cls._by_group[key] = pm
        return pm

    def __init__(self, device, group=None):
        self.device, self.group = device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._opened: Dict[tuple, int] = {}
        self.flag_pool = torch.zeros(FLAG_BLOCKS * FLAG_WORDS, dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)                     # zeros are in memory before any peer writes a flag
        self._flag_bases = self.map(self.flag_pool)
        self._next_block = 0

    def map(self, t: torch.Tensor) -> List[int]:
        """Addresses of tensor `t` (same role on every rank, collective call) in this process."""
        lib = _lib.load()
        handle = (ctypes.c_uint8 * 64)()
        off = ctypes.c_int64(0)
        _lib.check(lib.gae_ipc_get_handle(ctypes.c_void_p(t.data_ptr()), ctypes.byref(handle), ctypes.byref(off)),
                   "gae_ipc_get_handle")
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (bytes(handle), int(off.value)), group=self.group)
        ptrs = []
        for r, (h, o) in enumerate(everyone):
            if r == self.rank:
                ptrs.append(t.data_ptr())
                continue
            base = self._opened.get((r, h))
            if base is None:
                buf = (ctypes.c_uint8 * 64).from_buffer_copy(h)
                out = ctypes.c_void_p()
                _lib.check(lib.gae_ipc_open_handle(ctypes.byref(buf), ctypes.byref(out)), "gae_ipc_open_handle")
                base = int(out.value)
                self._opened[(r, h)] = base
            ptrs.append(base + o)
        return ptrs

    def flag_block(self):
        """(my block as an int64 view, its address on every rank); collective: every rank takes the
        same block index."""
        k = self._next_block
        if k >= FLAG_BLOCKS:
            raise GaeError("PeerMemory: out of flag blocks")
        self._next_block += 1
        mine = self.flag_pool[k * FLAG_WORDS:(k + 1) * FLAG_WORDS]
        return mine, [b + k * FLAG_WORDS * 8 for b in self._flag_bases]


class HaloSpMM:
    """Y_local = (A X)[my rows]: staged one-sided halo push overlapped with the row-block SpMMs, the
    whole operator behind one C call (gae_halo_spmm_f32).  `X_ext` is the [n_local + n_halo, d]
    feature buffer; callers write their rows into `X_local` on the current stream."""

    exchange = "halo"

    def __init__(self, hp: HaloPlan, d: int, n_stages: int = DEFAULT_STAGES, group=None, push_ctas: int = 0,
                 push_threads: int = 0, timeout_ms: int = 0, two_streams: bool = True, kind: str = "fold",
                 thresholds=None):
        if d % 4 != 0:
            raise GaeError("the halo exchange needs 16-byte rows (d a multiple of 4)")
        self.hp, self.d, self.group = hp, d, group
        dev = hp.rowptr.device
        if kind == "fold":
            hot, fold = tuple(thresholds) if thresholds else DEFAULT_FOLD
            self.sp = sp = build_fold_plan(hp, hot, fold, group)
        elif kind == "classes":
            self.sp = sp = build_class_plan(hp, thresholds or DEFAULT_HOT_THRESHOLDS, group)
        elif kind == "blocks":
            self.sp = sp = build_stage_plan(hp, n_stages, group)
        else:
            raise GaeError(f"unknown stage kind '{kind}'")
        self.n_ext = sp.n_ext if sp.n_ext >= 0 else hp.n_halo          # halo rows held in the buffer
        self.X_ext = ops.alloc_rows(hp.n_local + self.n_ext + sp.n_pre, d, dev)
        self.Y = ops.alloc_rows(hp.n_local, d, dev)
        # one segment workspace, a private slice per row block (consecutive blocks overlap in time)
        seg_counts = [p.n_seg if p is not None else 0 for p in sp.sub_plan]
        self.ws = torch.empty((max(sum(seg_counts), 1), ops.round_up4(d)), dtype=torch.float32, device=dev)
        seg_first = [sum(seg_counts[:s]) for s in range(sp.n_stages)]
        self.epoch = 0
        # high priority: the push kernels and the sums owed to the peers must get SMs ahead of the bulk
        # aggregation queued on the compute stream (measured on 2 GPUs: without it the folded rows of a rank
        # whose piece 0 started early were summed only after that piece, 2.1 ms late for its peers)
        self._comm = torch.cuda.Stream(device=dev, priority=-1)
        self._aux = torch.cuda.Stream(device=dev, priority=-1) if two_streams else None
        ex = HaloExchangeStruct()
        ex.world, ex.rank, ex.n_stages, ex.d = hp.world, hp.rank, sp.n_stages, d
        ex.ld = self.X_ext.stride(0)
        ex.x_local = self.X_ext.data_ptr()
        if hp.world > 1:
            pm = PeerMemory.get(dev, group)
            lds = [None] * hp.world
            dist.all_gather_object(lds, int(self.X_ext.stride(0)), group=group)
            if any(x != lds[0] for x in lds):
                raise GaeError("halo exchange needs the same feature row stride on every rank")
            self._peer_x = torch.tensor(pm.map(self.X_ext), dtype=torch.int64, device=dev)
            self.flags, flag_ptrs = pm.flag_block()
            self._peer_flags = torch.tensor(flag_ptrs, dtype=torch.int64, device=dev)
        else:
            self._peer_x = torch.tensor([self.X_ext.data_ptr()], dtype=torch.int64, device=dev)
            self.flags = torch.zeros(FLAG_WORDS, dtype=torch.int64, device=dev)
            self._peer_flags = torch.tensor([self.flags.data_ptr()], dtype=torch.int64, device=dev)
        ex.peer_x, ex.peer_flags, ex.flags = self._peer_x.data_ptr(), self._peer_flags.data_ptr(), self.flags.data_ptr()
        ex.send_src, ex.send_peer, ex.send_dst = sp.push_src.data_ptr(), sp.push_peer.data_ptr(), sp.push_dst.data_ptr()
        ex.stage_ptr = sp.stage_ptr.ctypes.data
        self._stage_done = torch.zeros(sp.n_stages, dtype=torch.int32, device=dev)
        ex.stage_done = self._stage_done.data_ptr()
        ex.push_ctas, ex.push_threads, ex.timeout_ms = int(push_ctas), int(push_threads), int(timeout_ms)
        if sp.n_pre > 0:
            ex.pre_rowptr, ex.pre_col = sp.pre_rowptr.data_ptr(), sp.pre_col.data_ptr()
            ex.pre_n_rows, ex.pre_row0, ex.pre_stage = sp.n_pre, hp.n_local + self.n_ext, sp.pre_stage
            pp = sp.pre_plan
            if pp is not None and (pp.n_seg > 0 or pp.bins is not None):
                ex.pre_plan = ctypes.addressof(pp.struct)
            self._pre_ws = pp.workspace(d, dev) if pp is not None and pp.n_seg > 0 else None
            ex.pre_ws = self._pre_ws.data_ptr() if self._pre_ws is not None else None
        self._ex = ex
        blocks = (HaloBlockStruct * sp.n_stages)()
        for s in range(sp.n_stages):
            blocks[s].row0, blocks[s].n_rows = sp.piece_rows[s]
            blocks[s].accumulate = int(sp.accumulate[s])
            blocks[s].rowptr, blocks[s].col = sp.sub_rowptr[s].data_ptr(), sp.sub_col[s].data_ptr()
            p = sp.sub_plan[s]
            if p is not None and (p.n_seg > 0 or p.bins is not None):
                blocks[s].plan = ctypes.addressof(p.struct)
            blocks[s].partial_ws = self.ws[seg_first[s]:].data_ptr() if seg_counts[s] else None
        self._blocks = blocks
        if hp.world > 1:
            torch.cuda.synchronize(dev)
            dist.barrier(group=group)      # every rank's buffers and zeroed flags exist before the first push

    @property
    def X_local(self) -> torch.Tensor:
        return self.X_ext[: self.hp.n_local]

    @property
    def X_halo(self) -> torch.Tensor:
        """Halo rows in the order of HaloPlan.halo_ids (a gathered copy when the buffer is laid out by class)."""
        if self.sp.kind == "blocks":
            return self.X_ext[self.hp.n_local:]
        if self.sp.kind == "fold":
            raise GaeError("kind 'fold' keeps only part of the halo rows (the rest arrives folded)")
        return self.X_ext[self.sp.halo_pos.to(self.X_ext.device)]

    def __call__(self) -> torch.Tensor:
        self.epoch += 1
        rc = _lib.load().gae_halo_spmm_f32(ctypes.byref(self._ex), self._blocks, ctypes.c_void_p(self.Y.data_ptr()),
                                           self.Y.stride(0), self.epoch, ops._stream(), self._comm.cuda_stream,
                                           self._aux.cuda_stream if self._aux is not None else None)
        _lib.check(rc, "gae_halo_spmm_f32")
        return self.Y

    def trace(self) -> dict:
        """Timeline of the last call in microseconds after the push started (synchronises): when the push
        published each stage to the peers, and when the consumer began / stopped waiting for each stage."""
        torch.cuda.synchronize(self.flags.device)
        off = 16 * 32 + 16 + 1                       # HALO_TRACE_OFF
        w = self.flags.cpu().numpy().astype(np.int64)
        t0 = int(w[off])
        us = lambda x: round((int(x) - t0) / 1e3, 1)  # noqa: E731
        n = self.sp.n_stages
        return {"published_us": [us(w[off + 2 + 3 * s]) for s in range(n)],
                "wait_begin_us": [us(w[off + 2 + 3 * s + 1]) for s in range(n)],
                "wait_end_us": [us(w[off + 2 + 3 * s + 2]) for s in range(n)],
                "release_us": us(w[off + 1]),
                "stage_rows": [int(self.sp.stage_ptr[s + 1] - self.sp.stage_ptr[s]) for s in range(n)]}

    def check(self) -> None:
        """Synchronous: raise if any flag wait of this operator timed out."""
        t = ctypes.c_int64(0)
        _lib.check(_lib.load().gae_halo_status(ctypes.byref(self._ex), ctypes.byref(t)), "gae_halo_status")


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
        hp = self.hp
        lds = [None] * hp.world
        dist.all_gather_object(lds, int(self.X_ext.stride(0)), group=self.group)
        if any(x != lds[0] for x in lds):
            raise GaeError("p2p exchange needs the same feature row stride on every rank")
        ptrs = PeerMemory.get(self.X_ext.device, self.group).map(self.X_ext)
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


def _halo_desc(op: "HaloSpMM") -> str:
    sp = op.sp
    if sp.kind == "fold":
        st = sp.stats
        return (f"stages: remote sources referenced >= {st['hot']} times are copied first, the rest of the remote edges "
                f"arrive as rows FOLDED by their owner (one sum per (owner, destination row) with >= {st['fold']} such "
                f"edges) or as cold copies; rank 0 receives {st['rows_in']} rows instead of {st['halo_rows_dedup']} "
                "deduplicated halo rows; the aggregation over local + hot sources overlaps the transfer of the rest")
    if sp.kind == "classes":
        return (f"{sp.n_stages} popularity classes of deduplicated halo rows (hot rows first; the aggregation over local + "
                "hot sources overlaps the transfer of the rarely used rows, which are then added)")
    return f"{sp.n_stages} row-block stages of deduplicated halo rows overlapped with the row-block SpMMs"


def build_rmat_partition(scale: int, total_edges: int, seed: int, d: int, device, exchange: str = "auto",
                         group=None, stages: int = DEFAULT_STAGES, push_ctas: int = 0,
                         two_streams: bool = True, kind: str = "fold", thresholds=None) -> RmatPartition:
    """Distributed R-MAT workload: every rank draws 1/P of the edge stream, edges are routed to the owner
    of their row, and the forward (rows = dst) and backward (rows = src) partitioned operators are built.
    exchange: "halo" (staged one-sided push with device flags, overlapped with `stages` row blocks),
    "nccl", "push", "p2p"; "auto" = halo, falling back to nccl collectively when CUDA IPC is unavailable."""
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
        exchange = "halo"
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
        if mode == "halo":
            kw = dict(push_ctas=push_ctas, two_streams=two_streams, kind=kind, thresholds=thresholds)
            return HaloSpMM(hp_f, d, stages, group, **kw), HaloSpMM(hp_b, d, stages, group, **kw)
        return PartitionedSpMM(hp_f, d, mode, group), PartitionedSpMM(hp_b, d, mode, group)

    if requested == "auto":
        # CUDA IPC needs peer access between every pair of GPUs; agree collectively, else use NCCL
        err = None
        try:
            fwd_op, bwd_op = make_ops("halo")
            ok = 1
        except GaeError as exc:
            fwd_op = bwd_op = None
            ok, err = 0, exc
        flag = torch.tensor([ok], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag) == 0:
            if rank == 0:
                print(f"[gae_dgl_b200.parallel] one-sided halo exchange unavailable ({err}); using NCCL all-to-all-v")
            exchange = "nccl"
            del fwd_op, bwd_op
            fwd_op, bwd_op = make_ops("nccl")
    else:
        fwd_op, bwd_op = make_ops(exchange)
    lo = bounds[rank]
    fwd_op.X_local.copy_(synthetic.hashed_normal(hp_f.n_local, d, 2, device=device, first_row=lo))
    bwd_op.X_local.copy_(synthetic.hashed_normal(hp_b.n_local, d, 3, device=device, first_row=lo))
    desc = {"nccl": "pack + NCCL all-to-all-v of deduplicated halo rows, per SpMM",
            "push": "one-sided push of deduplicated halo rows into peer HBM (CUDA IPC, posted NVLink stores) between two NCCL barriers, per SpMM",
            "p2p": "one-sided pull of deduplicated halo rows from peer HBM (CUDA IPC over NVLink), per SpMM",
            "halo": "one-sided staged push into peer HBM (CUDA IPC, posted NVLink stores, device-side release/acquire "
                    "flags, no collective), per SpMM; " + (_halo_desc(fwd_op) if exchange == "halo" else "")}[exchange]
    return RmatPartition(fwd_op, bwd_op, hp_f.n_edges, hp_f.n_local, hp_f.n_halo, desc, total_edges, d)
