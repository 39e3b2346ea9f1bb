#!/usr/bin/env python
"""Timeline of one partitioned SpMM (exchange 'halo') per rank: run under torch.distributed.run.
    python -m torch.distributed.run --nproc-per-node N tools/halo_trace.py [--stages B] [--push-ctas C] [--scale S --edges E]"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stages", type=int, default=8)
    ap.add_argument("--push-ctas", type=int, default=0)
    ap.add_argument("--one-stream", action="store_true")
    ap.add_argument("--push-threads", type=int, default=0)
    ap.add_argument("--kind", type=str, default="fold")
    ap.add_argument("--tune", type=str, default="")
    ap.add_argument("--hot", type=str, default="")
    ap.add_argument("--base-scale", type=int, default=22)
    ap.add_argument("--base-edges", type=int, default=100_000_000)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from gae_dgl_b200 import _lib as L, ops, parallel
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        L.set_tuning(k, int(v))
    scale = args.base_scale + int(round(math.log2(world)))
    part = parallel.build_rmat_partition(scale, args.base_edges * world, seed=1, d=64, device=dev, exchange="halo",
                                         stages=args.stages, push_ctas=args.push_ctas, two_streams=not args.one_stream, kind=args.kind,
                                         thresholds=tuple(int(x) for x in args.hot.split(",")) if args.hot else None)
    if args.push_threads:
        part.fwd_op._ex.push_threads = part.bwd_op._ex.push_threads = args.push_threads
    for _ in range(5):
        part.fwd()
        part.bwd()
    torch.cuda.synchronize()
    dist.barrier()
    out = {"rank": rank, "stages": part.fwd_op.sp.n_stages, "kind": part.fwd_op.sp.kind,
           "piece_edges": [int(r[-1]) for r in part.fwd_op.sp.sub_rowptr], "fwd": part.fwd_op.trace(), "bwd": part.bwd_op.trace()}
    # the pieces on their own: the push alone (no consumer), and the row-block SpMMs alone (no waits)
    op = part.fwd_op
    st = torch.cuda.current_stream()

    def timed(fn, k=5):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(k):
            fn()
        e1.record(st)
        e1.synchronize()
        return round(e0.elapsed_time(e1) / k, 3)

    sp = op.sp

    def blocks_only():
        for s in range(sp.n_stages):
            r0, nr = sp.piece_rows[s]
            if nr > 0:
                ops.spmm(sp.sub_rowptr[s], sp.sub_col[s], op.X_ext, sp.sub_plan[s], out=op.Y[r0:r0 + nr],
                         accumulate=sp.accumulate[s],
                         partial_ws=op.ws if sp.sub_plan[s] is not None and sp.sub_plan[s].n_seg else None)

    out["blocks_only_ms"] = timed(blocks_only)
    xfull = ops.alloc_rows(op.hp.n_local + op.hp.n_halo, 64, dev).normal_()     # the unfolded [local | halo] buffer
    wsf = op.hp.plan.workspace(64, dev)
    out["whole_plan_spmm_ms"] = timed(lambda: ops.spmm(op.hp.rowptr, op.hp.col, xfull, op.hp.plan, out=op.Y, partial_ws=wsf))
    del xfull
    if sp.n_pre:
        out["exchange"] = sp.stats
        pre_out = op.X_ext[op.hp.n_local + op.n_ext:]
        out["fold_spmm_only_ms"] = timed(lambda: ops.spmm(sp.pre_rowptr, sp.pre_col, op.X_ext, sp.pre_plan, out=pre_out,
                                                          partial_ws=op._pre_ws))
    out["op_ms"] = timed(lambda: op())
    op.check()
    # the push alone: nothing else runs on the GPU; every stage is waited for, then the halo is released
    import ctypes
    from gae_dgl_b200 import _lib
    lib = _lib.load()

    def push_only():
        op.epoch += 1
        _lib.check(lib.gae_halo_push_f32(ctypes.byref(op._ex), op.epoch, op._comm.cuda_stream), "push")   # staging rows as they are
        for s in range(sp.n_stages):
            _lib.check(lib.gae_halo_wait_f32(ctypes.byref(op._ex), s, op.epoch, st.cuda_stream), "wait")
        _lib.check(lib.gae_halo_release_f32(ctypes.byref(op._ex), op.epoch, st.cuda_stream), "release")
        op._comm.wait_stream(st)

    op._comm.wait_stream(st)
    out["push_only_ms"] = timed(push_only)
    out["push_only"] = op.trace()
    op.check()
    for r in range(world):
        if r == rank:
            print(json.dumps(out), flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
