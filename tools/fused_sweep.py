#!/usr/bin/env python
"""Separate launches vs the single-launch SpMM forward (tuning spmm_fused) over graph sizes, bit-identity checked."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gae_dgl_b200 import _lib, ops, synthetic  # noqa: E402
from gae_dgl_b200.graph import coo_to_csr_torch  # noqa: E402


def time_ms(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    for scale, edges, d in ((14, 400_000, 64), (16, 1_500_000, 64), (18, 6_000_000, 64), (19, 12_500_000, 64),
                            (20, 25_000_000, 64), (22, 100_000_000, 64), (22, 100_000_000, 32), (19, 12_500_000, 16)):
        n = 1 << scale
        src, dst = synthetic.rmat_edges(scale, edges, seed=1, device=dev)
        rowptr, col = coo_to_csr_torch(src, dst, n)
        del src, dst
        plan = ops.build_hub_plan(rowptr, 512, bins=True)
        ops.order_segments_by_source(plan, rowptr, col)
        X = synthetic.hashed_normal(n, d, 2, device=dev)
        ws = plan.workspace(d, dev)
        Y0, Y1 = torch.empty_like(X), torch.full_like(X, float("nan"))
        res = {"scale": scale, "edges": edges, "d": d}
        for knob, Y in ((0, Y0), (1, Y1)):
            _lib.set_tuning("spmm_fused", knob)
            ops.spmm(rowptr, col, X, plan, out=Y, partial_ws=ws)
            res["fused_ms" if knob else "separate_ms"] = time_ms(lambda: ops.spmm(rowptr, col, X, plan, out=Y, partial_ws=ws),
                                                                 20 if edges < 50_000_000 else 10)
        _lib.set_tuning("spmm_fused", -1)
        res["bit_identical"] = bool(torch.equal(Y0, Y1))
        print(json.dumps(res), flush=True)
        del rowptr, col, X, Y0, Y1, ws, plan
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
