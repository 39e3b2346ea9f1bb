"""Tensor-level wrappers over the C ABI + the autograd Functions of the hot path.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every arithmetic
kernel is in libgae_b200.so.  All entry points require CUDA float32 tensors and raise
otherwise -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import GaeError, HubPlanStruct

ACT_IDENTITY = 0
ACT_RELU = 1
DEC_LOSS = 1
DEC_GRAD = 2
DEFAULT_SEG_LEN = 512
MAX_FUSED_DECODER_WIDTH = 64      # dec_config (csrc/decoder.cu): wider embeddings take the materialised path


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------

_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device.  The raw getter costs ~1 us;
    torch.cuda.current_stream() builds a Stream object (~6 us, three times per inductive step)."""
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _require_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise GaeError(f"{name} must be a CUDA tensor (gae_dgl_b200 has no CPU path); got "
                       f"{type(t).__name__} on {getattr(t, 'device', None)}")
    if t.dtype != dtype:
        raise GaeError(f"{name} must be {dtype}, got {t.dtype}")
    # the library launches on the CURRENT device and stream: a tensor that lives elsewhere would be an
    # illegal access (or, with peer access enabled, a silent race against that device's stream)
    cur = _raw_device() if _raw_device is not None else torch.cuda.current_device()
    if t.device.index != cur:
        raise GaeError(f"{name} lives on {t.device} but the current CUDA device is cuda:{cur}; call "
                       f"torch.cuda.set_device({t.device.index}) (the trainers do) or wrap the call in torch.cuda.device(...)")


def round_up4(d: int) -> int:
    return (d + 3) // 4 * 4


def alloc_rows(n: int, d: int, device, zero: bool = False) -> torch.Tensor:
    """[n, d] fp32 view over a buffer whose row stride is a multiple of 4 floats, so the
    128-bit kernels apply to any d (d = 39 -> stride 40).  The padding
    columns are zero."""
    ld = round_up4(max(d, 1))
    if ld == d:
        return (torch.zeros if zero else torch.empty)((n, d), dtype=torch.float32, device=device)
    buf = torch.zeros((n, ld), dtype=torch.float32, device=device)
    return buf[:, :d]


def as_rows(t: torch.Tensor, name: str = "tensor") -> torch.Tensor:
    """Row-major [n, d] with unit column stride and a 16-byte aligned, multiple-of-4 row
    stride (copying into a padded buffer only when needed)."""
    _require_cuda(t, name)
    if t.dim() != 2:
        raise GaeError(f"{name} must be 2-D, got shape {tuple(t.shape)}")
    n, d = t.shape
    col_ok = t.stride(1) == 1 or d == 1
    if n > 1:
        row_ok = t.stride(0) >= round_up4(d) and t.stride(0) % 4 == 0
    else:
        row_ok = d % 4 == 0
    if col_ok and row_ok and t.data_ptr() % 16 == 0:
        return t
    out = alloc_rows(n, d, t.device)
    out.copy_(t)
    return out


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else round_up4(t.shape[1])


# ------------------------------------------------------------------------------------------
# hub plan
# ------------------------------------------------------------------------------------------

SHORT_MAX = 4
BIN_MIN_ROWS = 1 << 16     # below this a single row pass wins (fewer launches)


@dataclass
class HubPlan:
    seg_len: int
    n_long: int
    n_seg: int
    long_row: torch.Tensor
    long_seg_ptr: torch.Tensor
    seg_row: torch.Tensor
    struct: HubPlanStruct
    bins: Optional[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = None   # (empty, short, mid) row lists
    seg_order: Optional[torch.Tensor] = None                                 # int32 permutation of the segments

    def workspace(self, d: int, device) -> Optional[torch.Tensor]:
        if self.n_seg == 0:
            return None
        return torch.empty((self.n_seg, round_up4(d)), dtype=torch.float32, device=device)


def order_segments_by_source(plan: "HubPlan", rowptr: torch.Tensor, col: torch.Tensor) -> None:
    """Fill plan.seg_order: the hub segments sorted by their first source id (stable), so that
    warps running at the same time gather from the same band of X (L2-resident).  Index-only work
    done once per graph with device tensor ops."""
    if plan.n_seg == 0:
        return
    dev = rowptr.device
    seg = torch.arange(plan.n_seg, device=dev, dtype=torch.int64)
    k = plan.seg_row[:plan.n_seg].to(torch.int64)
    row = plan.long_row[:plan.n_long].to(torch.int64)[k]
    first_edge = rowptr[row] + (seg - plan.long_seg_ptr[k]) * plan.seg_len
    order = torch.argsort(col[first_edge].to(torch.int64), stable=True).to(torch.int32).contiguous()
    plan.seg_order = order
    plan.struct.seg_order = order.data_ptr()


MID_SORT_DEFAULT = True    # order the mid-row list by descending degree: the two warps of a CTA finish together (-3.6 % on C4)


def build_hub_plan(rowptr: torch.Tensor, seg_len: int = DEFAULT_SEG_LEN, bins: Optional[bool] = None,
                   sort_mid: Optional[bool] = None) -> HubPlan:
    """Split rows with in-degree > seg_len into fixed-length segments (gae_hub_plan_*_host) and,
    for large graphs, bin the remaining rows by degree (gae_row_bins_host).  Runs once per graph
    on the host copy of rowptr.  sort_mid orders the mid-row list by descending in-degree (stable),
    so that the two warps of a CTA (consecutive list entries) finish together; results do not depend on it."""
    lib = _lib.load()
    rp = rowptr.detach().to("cpu", torch.int64).contiguous().numpy()
    n_rows = rp.shape[0] - 1
    n_long, n_seg = ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check(lib.gae_hub_plan_count_host(rp.ctypes.data, n_rows, seg_len, ctypes.byref(n_long),
                                           ctypes.byref(n_seg)), "gae_hub_plan_count_host")
    nl, ns = n_long.value, n_seg.value
    long_row = np.zeros(max(nl, 1), dtype=np.int32)
    long_seg_ptr = np.zeros(nl + 1, dtype=np.int64)
    seg_row = np.zeros(max(ns, 1), dtype=np.int32)
    _lib.check(lib.gae_hub_plan_fill_host(rp.ctypes.data, n_rows, seg_len, long_row.ctypes.data,
                                          long_seg_ptr.ctypes.data, seg_row.ctypes.data), "gae_hub_plan_fill_host")
    dev = rowptr.device
    t_long = torch.from_numpy(long_row).to(dev)
    t_ptr = torch.from_numpy(long_seg_ptr).to(dev)
    t_seg = torch.from_numpy(seg_row).to(dev)
    st = HubPlanStruct(seg_len=seg_len, short_max=SHORT_MAX, n_long=nl, n_seg=ns, long_row=t_long.data_ptr(),
                       long_seg_ptr=t_ptr.data_ptr(), seg_row=t_seg.data_ptr())
    plan = HubPlan(seg_len, nl, ns, t_long, t_ptr, t_seg, st)
    if bins is None:
        bins = n_rows >= BIN_MIN_ROWS
    if bins:
        counts = (ctypes.c_int64 * 3)()
        _lib.check(lib.gae_row_bins_host(rp.ctypes.data, n_rows, seg_len, SHORT_MAX, ctypes.byref(counts), None, None,
                                         None), "gae_row_bins_host")
        arrs = [np.zeros(max(int(c), 1), dtype=np.int32) for c in counts]
        _lib.check(lib.gae_row_bins_host(rp.ctypes.data, n_rows, seg_len, SHORT_MAX, ctypes.byref(counts),
                                         arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data),
                   "gae_row_bins_host")
        ts = tuple(torch.from_numpy(a).to(dev) for a in arrs)
        n_mid = int(counts[2])
        if (MID_SORT_DEFAULT if sort_mid is None else sort_mid) and n_mid > 1:
            mid = ts[2][:n_mid].to(torch.int64)
            deg = rowptr[mid + 1] - rowptr[mid]
            ts[2][:n_mid] = mid[torch.argsort(deg, descending=True, stable=True)].to(torch.int32)
        plan.bins = ts
        st.n_empty, st.n_short, st.n_mid = int(counts[0]), int(counts[1]), int(counts[2])
        st.empty_rows, st.short_rows, st.mid_rows = (t.data_ptr() for t in ts)
    return plan


# ------------------------------------------------------------------------------------------
# raw ops
# ------------------------------------------------------------------------------------------

def spmm(rowptr: torch.Tensor, col: torch.Tensor, X: torch.Tensor, plan: Optional[HubPlan] = None,
         out: Optional[torch.Tensor] = None, accumulate: bool = False, vals: Optional[torch.Tensor] = None,
         partial_ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Y = A X  (gae_spmm_csr_f32).  rowptr int64 [n+1], col int32 [E] on the same device."""
    _require_cuda(rowptr, "rowptr", torch.int64)
    _require_cuda(col, "col", torch.int32)
    X = as_rows(X, "X")
    n_rows = rowptr.numel() - 1
    d = X.shape[1]
    if out is None:
        if accumulate:
            raise GaeError("accumulate=True needs an `out` tensor")
        out = alloc_rows(n_rows, d, X.device)
    else:
        _require_cuda(out, "out")
        if out.shape != (n_rows, d) or (out.stride(1) != 1 and d > 1):
            raise GaeError("out has the wrong shape / layout")
    plan_ref = None
    if plan is not None and (plan.n_seg > 0 or plan.bins is not None):
        plan_ref = ctypes.byref(plan.struct)
        if partial_ws is None:
            partial_ws = plan.workspace(d, X.device)
    if vals is not None:
        _require_cuda(vals, "vals")
    rc = _lib.load().gae_spmm_csr_f32(_ptr(rowptr), _ptr(col), _ptr(vals), _ptr(X), _ld(X), _ptr(out), _ld(out),
                                      n_rows, d, plan_ref, _ptr(partial_ws), int(accumulate), _stream())
    _lib.check(rc, "gae_spmm_csr_f32")
    return out


def linear_fwd(Y: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor], act: int) -> torch.Tensor:
    Y = as_rows(Y, "Y")
    _require_cuda(W, "W")
    W = W.contiguous()
    n, d_in = Y.shape
    d_out = W.shape[0]
    if W.shape[1] != d_in:
        raise GaeError(f"weight shape {tuple(W.shape)} does not match input width {d_in}")
    H = alloc_rows(n, d_out, Y.device)
    bb = None if b is None else b.contiguous()
    rc = _lib.load().gae_linear_fwd_f32(_ptr(Y), _ld(Y), _ptr(W), _ptr(bb), _ptr(H), _ld(H), n, d_in, d_out, act,
                                        _stream())
    _lib.check(rc, "gae_linear_fwd_f32")
    return H


def linear_bwd(Y: torch.Tensor, W: torch.Tensor, H: torch.Tensor, dH: torch.Tensor, act: int, need_dy: bool):
    Y = as_rows(Y, "Y")
    dH = as_rows(dH, "dH")
    W = W.contiguous()
    n, d_in = Y.shape
    d_out = W.shape[0]
    lib = _lib.load()
    ws_bytes = lib.gae_linear_bwd_ws_bytes(n, d_in, d_out)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=Y.device)
    dW = torch.empty((d_out, d_in), dtype=torch.float32, device=Y.device)
    db = torch.empty((d_out,), dtype=torch.float32, device=Y.device)
    dY = alloc_rows(n, d_in, Y.device) if need_dy else None
    rc = lib.gae_linear_bwd_f32(_ptr(Y), _ld(Y), _ptr(W), _ptr(H), _ld(H) if H is not None else 0, _ptr(dH), _ld(dH),
                                _ptr(dY), _ld(dY) if dY is not None else 0, _ptr(dW), _ptr(db), _ptr(ws), ws_bytes,
                                n, d_in, d_out, act, _stream())
    _lib.check(rc, "gae_linear_bwd_f32")
    return dY, dW, db


def gcn_layer_fwd(rowptr: torch.Tensor, col: torch.Tensor, Hin: torch.Tensor, W: torch.Tensor,
                  b: Optional[torch.Tensor], act: int, want_y: bool = False):
    """One GCN layer in one launch (gae.py:26-31): returns (act((A Hin) W^T + b), A Hin or None).
    d_in, d_out <= 64 (gae_gcn_layer_fwd_f32)."""
    _require_cuda(rowptr, "rowptr", torch.int64)
    _require_cuda(col, "col", torch.int32)
    Hin = as_rows(Hin, "Hin")
    _require_cuda(W, "W")
    W = W.contiguous()
    n, d_in = Hin.shape
    d_out = W.shape[0]
    if W.shape[1] != d_in or rowptr.numel() - 1 != n:
        raise GaeError("gcn_layer_fwd: shapes do not match")
    H = alloc_rows(n, d_out, Hin.device)
    Y = alloc_rows(n, d_in, Hin.device) if want_y else None
    bb = None if b is None else b.contiguous()
    rc = _lib.load().gae_gcn_layer_fwd_f32(_ptr(rowptr), _ptr(col), _ptr(Hin), _ld(Hin), _ptr(W), _ptr(bb), _ptr(H), _ld(H),
                                           _ptr(Y), _ld(Y) if Y is not None else 0, n, d_in, d_out, act, _stream())
    _lib.check(rc, "gae_gcn_layer_fwd_f32")
    return H, Y


def gcn_layer_bwd(rowptr_t: Optional[torch.Tensor], col_t: Optional[torch.Tensor], Y: torch.Tensor, W: torch.Tensor,
                  H: torch.Tensor, dH: torch.Tensor, act: int, need_dhin: bool, plan_t: Optional[HubPlan] = None):
    """Adjoint of gcn_layer_fwd (gae_gcn_layer_bwd_f32): returns (dHin or None, dW, db)."""
    Y = as_rows(Y, "Y")
    dH = as_rows(dH, "dH")
    W = W.contiguous()
    n, d_in = Y.shape
    d_out = W.shape[0]
    lib = _lib.load()
    ws_bytes = lib.gae_gcn_layer_bwd_ws_bytes(n, d_in, d_out)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=Y.device)
    dW = torch.empty((d_out, d_in), dtype=torch.float32, device=Y.device)
    db = torch.empty((d_out,), dtype=torch.float32, device=Y.device)
    dY = alloc_rows(n, d_in, Y.device) if need_dhin else None
    dHin = alloc_rows(n, d_in, Y.device) if need_dhin else None
    plan_ref, hub_ws = None, None
    if need_dhin and plan_t is not None and (plan_t.n_seg > 0 or plan_t.bins is not None):
        plan_ref = ctypes.byref(plan_t.struct)
        hub_ws = plan_t.workspace(d_in, Y.device)
    rc = lib.gae_gcn_layer_bwd_f32(_ptr(rowptr_t), _ptr(col_t), plan_ref, _ptr(hub_ws), _ptr(Y), _ld(Y), _ptr(W), _ptr(H), _ld(H),
                                   _ptr(dH), _ld(dH), _ptr(dY), _ld(dY) if dY is not None else 0, _ptr(dHin),
                                   _ld(dHin) if dHin is not None else 0, _ptr(dW), _ptr(db), _ptr(ws), ws_bytes, n, d_in, d_out,
                                   act, _stream())
    _lib.check(rc, "gae_gcn_layer_bwd_f32")
    return dHin, dW, db


def dropout_fwd(Z: torch.Tensor, p: float, mask: Optional[torch.Tensor] = None,
                seed: int = 0, offset: int = 0, rng_state: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Zd = Z * keep / (1-p).  `mask` (uint8/bool keep-mask [n,d]) is READ when given; otherwise
    the mask is drawn from Philox -- from the device-resident state `rng_state` (int64[2], advanced
    on the stream: CUDA-graph safe) when given, else from the host-side (seed, offset)."""
    Z = as_rows(Z, "Z")
    n, d = Z.shape
    Zd = alloc_rows(n, d, Z.device)
    lib = _lib.load()
    if mask is None:
        m = torch.empty((n, d), dtype=torch.uint8, device=Z.device)
        if rng_state is not None:
            _require_cuda(rng_state, "rng_state", torch.int64)
            rc = lib.gae_dropout_fwd_devrng_f32(_ptr(Z), _ld(Z), _ptr(Zd), _ld(Zd), _ptr(m), n, d, float(p),
                                                _ptr(rng_state), _stream())
            _lib.check(rc, "gae_dropout_fwd_devrng_f32")
            return Zd, m
        mode = 0
    else:
        m = mask.to(device=Z.device, dtype=torch.uint8).contiguous()
        if m.shape != (n, d):
            raise GaeError("dropout mask shape mismatch")
        mode = 1
    rc = lib.gae_dropout_fwd_f32(_ptr(Z), _ld(Z), _ptr(Zd), _ld(Zd), _ptr(m), n, d, float(p),
                                 int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), mode, _stream())
    _lib.check(rc, "gae_dropout_fwd_f32")
    return Zd, m


def dropout_bwd(dZd: torch.Tensor, mask: torch.Tensor, p: float, grad_scale: Optional[torch.Tensor] = None):
    dZd = as_rows(dZd, "dZd")
    n, d = dZd.shape
    dZ = alloc_rows(n, d, dZd.device)
    gs = None
    if grad_scale is not None:
        gs = grad_scale.to(device=dZd.device, dtype=torch.float32).reshape(1).contiguous()
    rc = _lib.load().gae_dropout_bwd_f32(_ptr(dZd), _ld(dZd), _ptr(mask), _ptr(dZ), _ld(dZ), n, d, float(p), _ptr(gs),
                                         _stream())
    _lib.check(rc, "gae_dropout_bwd_f32")
    return dZ


def decoder_bce(Zd: torch.Tensor, rowptr, col, rowptr_t, col_t, pos_weight: float, want_loss=True,
                want_grad=False):
    """Fused decoder + BCE: returns (loss 0-dim tensor or None, dZd_unit or None)."""
    Zd = as_rows(Zd, "Zd")
    n, d = Zd.shape
    lib = _lib.load()
    ws_bytes = lib.gae_decoder_ws_bytes(n, d)
    if ws_bytes <= 0:
        raise GaeError(f"decoder does not support n={n}, d={d} (d must be <= 64)")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=Zd.device)
    loss = torch.empty((), dtype=torch.float32, device=Zd.device) if want_loss else None
    dZ = alloc_rows(n, d, Zd.device) if want_grad else None
    mode = (DEC_LOSS if want_loss else 0) | (DEC_GRAD if want_grad else 0)
    rc = lib.gae_decoder_bce_f32(_ptr(Zd), _ld(Zd), n, d, _ptr(rowptr), _ptr(col), _ptr(rowptr_t), _ptr(col_t),
                                 float(pos_weight), mode, _ptr(loss), _ptr(dZ), _ld(dZ) if dZ is not None else 0,
                                 _ptr(ws), ws_bytes, _stream())
    _lib.check(rc, "gae_decoder_bce_f32")
    return loss, dZ


def decoder_bce_blockdiag(Zd: torch.Tensor, rowptr, col, rowptr_t, col_t, blk_lo, blk_hi, n_pairs: float,
                          pos_weight: float, want_loss=True, want_grad=False):
    """Per-graph (block-diagonal) fused decoder + BCE for batched graphs."""
    Zd = as_rows(Zd, "Zd")
    n, d = Zd.shape
    lib = _lib.load()
    ws_bytes = lib.gae_decoder_blockdiag_ws_bytes(n, d)
    if ws_bytes <= 0:
        raise GaeError(f"decoder does not support n={n}, d={d} (d must be <= 64)")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=Zd.device)
    loss = torch.empty((), dtype=torch.float32, device=Zd.device) if want_loss else None
    dZ = alloc_rows(n, d, Zd.device) if want_grad else None
    mode = (DEC_LOSS if want_loss else 0) | (DEC_GRAD if want_grad else 0)
    rc = lib.gae_decoder_bce_blockdiag_f32(_ptr(Zd), _ld(Zd), n, d, _ptr(rowptr), _ptr(col), _ptr(rowptr_t), _ptr(col_t),
                                           _ptr(blk_lo), _ptr(blk_hi), float(n_pairs), float(pos_weight), mode,
                                           _ptr(loss), _ptr(dZ), _ld(dZ) if dZ is not None else 0, _ptr(ws), ws_bytes,
                                           _stream())
    _lib.check(rc, "gae_decoder_bce_blockdiag_f32")
    return loss, dZ


def decoder_logits(Zd: torch.Tensor) -> torch.Tensor:
    Zd = as_rows(Zd, "Zd")
    n, d = Zd.shape
    X = torch.empty((n, n), dtype=torch.float32, device=Zd.device)
    rc = _lib.load().gae_decoder_logits_f32(_ptr(Zd), _ld(Zd), n, d, _ptr(X), n, _stream())
    _lib.check(rc, "gae_decoder_logits_f32")
    return X


def in_degrees(rowptr: torch.Tensor) -> torch.Tensor:
    _require_cuda(rowptr, "rowptr", torch.int64)
    n = rowptr.numel() - 1
    deg = torch.empty(n, dtype=torch.int64, device=rowptr.device)
    _lib.check(_lib.load().gae_in_degrees_i64(_ptr(rowptr), n, _ptr(deg), _stream()), "gae_in_degrees_i64")
    return deg


def gather_rows(X: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    X = as_rows(X, "X")
    _require_cuda(idx, "idx", torch.int64)
    m, d = idx.numel(), X.shape[1]
    if out is None:
        out = alloc_rows(m, d, X.device)
    rc = _lib.load().gae_gather_rows_f32(_ptr(X), _ld(X), _ptr(idx), m, d, _ptr(out), _ld(out), _stream())
    _lib.check(rc, "gae_gather_rows_f32")
    return out


# ------------------------------------------------------------------------------------------
# autograd
# ------------------------------------------------------------------------------------------

class SpMMFunction(torch.autograd.Function):
    """update_all(copy_src, sum): Y = A X forward, dX = A^T dY backward (reference gae.py:28)."""

    @staticmethod
    def forward(ctx, X, graph):
        csr = graph.csr()
        ctx.graph = graph
        return spmm(csr.rowptr, csr.col, X, csr.plan)

    @staticmethod
    def backward(ctx, dY):
        if not ctx.needs_input_grad[0]:
            return None, None
        t = ctx.graph.csr_t()
        return spmm(t.rowptr, t.col, dY.contiguous(), t.plan), None


class LinearActFunction(torch.autograd.Function):
    """NodeApplyModule: act(Y W^T + b) (reference gae.py:13-16)."""

    @staticmethod
    def forward(ctx, Y, W, b, act):
        H = linear_fwd(Y, W, b, act)
        ctx.act = act
        ctx.has_bias = b is not None
        ctx.save_for_backward(Y, W, H)
        return H

    @staticmethod
    def backward(ctx, dH):
        Y, W, H = ctx.saved_tensors
        dY, dW, db = linear_bwd(Y, W, H, dH, ctx.act, need_dy=ctx.needs_input_grad[0])
        return dY, dW, (db if ctx.has_bias else None), None


class DecoderLossFunction(torch.autograd.Function):
    """dropout -> Zd Zd^T -> weighted BCE-with-logits (mean), fused; the gradient w.r.t. Zd is
    produced in the same pass as the loss and scaled by grad_output in backward."""

    @staticmethod
    def forward(ctx, Z, graph, pos_weight, p, mask, rng_state, per_graph=False):
        csr, csr_t = graph.csr(), graph.csr_t()
        need_grad = Z.requires_grad
        Zd, m = dropout_fwd(Z, p, mask, rng_state=rng_state)
        if per_graph:
            lo, hi, n_pairs = graph.block_ranges()
            loss, dZd_unit = decoder_bce_blockdiag(Zd, csr.rowptr, csr.col, csr_t.rowptr, csr_t.col, lo, hi, n_pairs,
                                                   pos_weight, want_loss=True, want_grad=need_grad)
        else:
            loss, dZd_unit = decoder_bce(Zd, csr.rowptr, csr.col, csr_t.rowptr, csr_t.col, pos_weight,
                                         want_loss=True, want_grad=need_grad)
        ctx.p = p
        ctx.mask = m
        ctx.dZd_unit = dZd_unit
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.dZd_unit is None:
            return (None,) * 7
        dZ = dropout_bwd(ctx.dZd_unit, ctx.mask, ctx.p, grad_scale=g)
        return dZ, None, None, None, None, None, None


class DecoderLogitsFunction(torch.autograd.Function):
    """Materialised logits (what gae.GAE.forward returns, gae.py:54-55).  Compatibility path:
    its backward uses torch.matmul on the N x N gradient the caller's dense loss produced."""

    @staticmethod
    def forward(ctx, Z, p, mask, rng_state):
        Zd, m = dropout_fwd(Z, p, mask, rng_state=rng_state)
        ctx.p = p
        ctx.save_for_backward(Zd, m)
        return decoder_logits(Zd)

    @staticmethod
    def backward(ctx, G):
        Zd, m = ctx.saved_tensors
        dZd = (G + G.t()) @ Zd
        return dropout_bwd(dZd, m, ctx.p), None, None, None


class FusedStepFunction(torch.autograd.Function):
    """Whole GAE step behind one C call (gae_step_fwd_bwd_f32): encoder forward, dropout, fused
    decoder loss and -- when any parameter needs a gradient -- the complete backward, computed
    eagerly for grad_output = 1 and scaled by the incoming gradient in backward().
    Returns (loss, embeddings); the embeddings are not differentiable through this path."""

    @staticmethod
    def forward(ctx, X, graph, dims, acts, pos_weight, p, mask, rng_state, per_graph, need_grad, *params):
        x_aggregated = False
        if isinstance(per_graph, tuple):              # (per_graph, x_aggregated): keeps the positional signature
            per_graph, x_aggregated = per_graph
        from ._lib import StepDesc
        lib = _lib.load()
        X = as_rows(X, "features")
        n = X.shape[0]
        L = len(dims) - 1
        csr, csr_t = graph.csr(), graph.csr_t()
        desc = StepDesc()
        desc.n_layers = L
        for i, v in enumerate(dims):
            desc.dims[i] = int(v)
        for i, a in enumerate(acts):
            desc.acts[i] = int(a)
        desc.dropout_p, desc.pos_weight, desc.per_graph = float(p), float(pos_weight), int(bool(per_graph))
        desc.x_aggregated = int(bool(x_aggregated))
        Ws = [params[2 * l].contiguous() for l in range(L)]
        bs = [params[2 * l + 1].contiguous() for l in range(L)]
        want_grad = bool(need_grad)
        dev = X.device
        plan = ctypes.byref(csr.plan.struct) if csr.plan is not None and (csr.plan.n_seg or csr.plan.bins) else None
        plan_t = ctypes.byref(csr_t.plan.struct) if csr_t.plan is not None and (csr_t.plan.n_seg or csr_t.plan.bins) else None
        ws_bytes = lib.gae_step_ws_bytes(ctypes.byref(desc), n, plan, plan_t)
        if ws_bytes <= 0:
            raise GaeError("gae_step_ws_bytes rejected the configuration")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        Z = alloc_rows(n, dims[-1], dev)
        dW = [torch.empty_like(w) for w in Ws] if want_grad else []
        db = [torch.empty_like(b) for b in bs] if want_grad else []
        arr = lambda ts: (ctypes.c_void_p * L)(*[t.data_ptr() for t in ts]) if ts else None  # noqa: E731
        lo = hi = None
        n_pairs = 0.0
        if per_graph:
            lo, hi, n_pairs = graph.block_ranges()
        m = None
        if mask is not None:
            m = mask.to(device=dev, dtype=torch.uint8).contiguous()
            if m.shape != (n, dims[-1]):
                raise GaeError("dropout mask shape mismatch")
        elif rng_state is None:
            raise GaeError("fused step needs a mask or a device RNG state")
        rc = lib.gae_step_fwd_bwd_f32(ctypes.byref(desc), n, _ptr(csr.rowptr), _ptr(csr.col), plan, _ptr(csr_t.rowptr),
                                      _ptr(csr_t.col), plan_t, _ptr(X), _ld(X), arr(Ws), arr(bs), _ptr(m),
                                      _ptr(rng_state), _ptr(lo), _ptr(hi), float(n_pairs), int(want_grad), _ptr(loss),
                                      _ptr(Z), _ld(Z), arr(dW), arr(db), _ptr(ws), ws_bytes, _stream())
        _lib.check(rc, "gae_step_fwd_bwd_f32")
        ctx.grads = [g for pair in zip(dW, db) for g in pair] if want_grad else None
        ctx.mark_non_differentiable(Z)
        return loss, Z

    @staticmethod
    def backward(ctx, g_loss, _g_z):
        if ctx.grads is None:
            return (None,) * 10
        # out of place: the cached unit gradients stay valid for a second backward (retain_graph) and never
        # alias what autograd installs as .grad
        return (None,) * 10 + tuple(torch._foreach_mul(ctx.grads, g_loss))
