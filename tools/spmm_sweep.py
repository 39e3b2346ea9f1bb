#!/usr/bin/env python
"""SpMM tuning sweep / profiling driver on the C4 workload (R-MAT scale 22, 1e8 edges, d=64).

    python tools/spmm_sweep.py --sweep            # time every tuning combination (CUDA events)
    python tools/spmm_sweep.py --iters 3          # a few launches with the current defaults (for ncu)
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gae_dgl_b200 import _lib, ops, synthetic  # noqa: E402
from gae_dgl_b200.graph import coo_to_csr_torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--edges", type=int, default=100_000_000)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--no-permute", action="store_true")
    ap.add_argument("--seg-lens", type=str, default="512")
    ap.add_argument("--tune", type=str, default="")
    ap.add_argument("--blocks", type=str, default="32,64,128,256")
    ap.add_argument("--caches", type=str, default="0,1")
    ap.add_argument("--variants", type=str, default="0")
    ap.add_argument("--unrolls", type=str, default="4,8")
    ap.add_argument("--bins", type=str, default="1")
    ap.add_argument("--stages", type=str, default="2,3,4")
    ap.add_argument("--out", type=str, default="")
    ap.add_argument("--mid-sort", type=str, default="0", help="0 / 1: degree-sorted mid-row list")
    ap.add_argument("--seg-orders", type=str, default="0,1")
    ap.add_argument("--fused", type=str, default="0", help="spmm_fused values (experimental one-launch forward)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n = 1 << args.scale
    src, dst = synthetic.rmat_edges(args.scale, args.edges, seed=1, device=dev, permute=not args.no_permute)
    rowptr, col = coo_to_csr_torch(src, dst, n)
    del src, dst
    torch.cuda.empty_cache()
    deg = rowptr[1:] - rowptr[:-1]
    print(f"graph: V={n} E={args.edges} max_indeg={int(deg.max())} empty_rows={float((deg == 0).float().mean()):.3f} "
          f"rows>512={int((deg > 512).sum())} edges_in_rows>512={int(deg[deg > 512].sum())}", flush=True)
    X = synthetic.hashed_normal(n, args.d, 2, device=dev)
    Y = torch.empty_like(X)
    alg = args.edges * (4 + 4 * args.d) + n * (4 * args.d + 4)
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        _lib.set_tuning(k, int(v))
    st = torch.cuda.current_stream()

    def timeit(plan, ws, iters):
        for _ in range(2):
            ops.spmm(rowptr, col, X, plan, out=Y, partial_ws=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(iters):
            ops.spmm(rowptr, col, X, plan, out=Y, partial_ws=ws)
        e1.record(st)
        e1.synchronize()
        return e0.elapsed_time(e1) / iters

    results = []
    blocks = [int(b) for b in args.blocks.split(',')]
    caches = [int(b) for b in args.caches.split(',')]
    unrolls = [int(b) for b in args.unrolls.split(',')]
    seg_lens = [int(s) for s in args.seg_lens.split(",")]
    if not args.sweep:
        plan = ops.build_hub_plan(rowptr, seg_lens[0])
        ops.order_segments_by_source(plan, rowptr, col)
        ws = plan.workspace(args.d, dev)
        ms = timeit(plan, ws, args.iters)
        print(f"default tuning: {ms:.3f} ms  {alg / ms / 1e6:.0f} GB/s algorithmic  {args.edges / ms / 1e6:.2f} Gedges/s")
        return
    ref = None
    for seg, mid_sort in itertools.product(seg_lens, [int(v) for v in args.mid_sort.split(",")]):
        plan = ops.build_hub_plan(rowptr, seg, sort_mid=bool(mid_sort))
        ops.order_segments_by_source(plan, rowptr, col)
        ws = plan.workspace(args.d, dev)
        for variant in [int(v) for v in args.variants.split(",")]:
            if variant == 0:
                continue
            _lib.set_tuning("spmm_variant", variant)
            for stages in [int(v) for v in args.stages.split(",")]:
                _lib.set_tuning("spmm_stages", stages)
                ms = timeit(plan, ws, args.iters)
                if ref is None:
                    ref = Y.clone()
                r = {"seg_len": seg, "variant": variant, "stages": stages, "ms": ms, "alg_GBps": alg / ms / 1e6,
                     "maxdiff_vs_first": float((Y - ref).abs().max())}
                results.append(r)
                print(json.dumps(r), flush=True)
        _lib.set_tuning("spmm_variant", 0)
        if "0" not in args.variants.split(","):
            continue
        for block, unroll, cache, rpw, bins, so, fused in itertools.product(
                blocks, unrolls, caches, (1,), [int(b) for b in args.bins.split(",")],
                [int(v) for v in args.seg_orders.split(",")], [int(v) for v in args.fused.split(",")]):
            _lib.set_tuning("spmm_fused", fused)
            _lib.set_tuning("spmm_bins", bins)
            _lib.set_tuning("spmm_seg_order", so)
            _lib.set_tuning("spmm_block", block)
            _lib.set_tuning("spmm_unroll", unroll)
            _lib.set_tuning("spmm_cache", cache)
            _lib.set_tuning("spmm_rows_per_warp", rpw)
            ms = timeit(plan, ws, args.iters)
            if ref is None:
                ref = Y.clone()
            err = float((Y - ref).abs().max())
            r = {"seg_len": seg, "mid_sort": mid_sort, "fused": fused, "block": block, "unroll": unroll, "cache": cache,
                 "rows_per_warp": rpw, "bins": bins, "seg_order": so, "ms": ms,
                 "alg_GBps": alg / ms / 1e6, "maxdiff_vs_first": err}
            results.append(r)
            print(json.dumps(r), flush=True)
    best = min(results, key=lambda r: r["ms"])
    print("BEST", json.dumps(best))
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"results": results, "best": best}, f, indent=1)


if __name__ == "__main__":
    main()
