#!/usr/bin/env python
"""Build the partitioned R-MAT workload ONCE and time the SpMM fwd+bwd step with every exchange
mechanism (run under torch.distributed.run)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--base-scale", type=int, default=22)
    ap.add_argument("--base-edges", type=int, default=100_000_000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--modes", type=str, default="push,nccl")
    ap.add_argument("--stages", type=str, default="", help="comma list, e.g. 2,4,8: also time parallel_staged on each mode")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from gae_dgl_b200 import parallel
    import math
    scale = args.base_scale + int(round(math.log2(world)))
    total = args.base_edges * world
    part = parallel.build_rmat_partition(scale, total, seed=1, d=64, device=dev, exchange="nccl")
    out = {"world": world, "scale": scale, "edges": total, "halo_rows_rank0": part.halo_rows}
    st = torch.cuda.current_stream()
    for mode in args.modes.split(","):
        ops_ = []
        for op in (part.fwd_op, part.bwd_op):
            new = parallel.PartitionedSpMM(op.hp, 64, mode)
            new.X_local.copy_(op.X_local)
            ops_.append(new)
        f, b = ops_
        for _ in range(3):
            f(); b()
        torch.cuda.synchronize(); dist.barrier()
        # exchange alone
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(st)
        for _ in range(args.steps):
            f.exchange_halo(); b.exchange_halo()
        e1.record(st)
        for _ in range(args.steps):
            f(); b()
        e2.record(st)
        e2.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps, e1.elapsed_time(e2) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[mode] = {"exchange_only_ms_per_step": float(t[0]), "step_ms": float(t[1]),
                     "edges_per_s": total / (float(t[1]) * 1e-3)}
        # staged exchange pipelined with row-block SpMMs, same partition, same buffers
        for n_stages in [int(v) for v in args.stages.split(",") if v]:
            if mode not in ("push", "nccl"):
                continue
            from gae_dgl_b200.parallel_staged import StagedPartitionedSpMM
            ref_f, ref_b = f().clone(), b().clone()
            sf, sb = StagedPartitionedSpMM(f, n_stages), StagedPartitionedSpMM(b, n_stages)
            for _ in range(3):
                sf(); sb()
            torch.cuda.synchronize(); dist.barrier()
            err = max(float((sf() - ref_f).abs().max()), float((sb() - ref_b).abs().max()))
            torch.cuda.synchronize(); dist.barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(st)
            for _ in range(args.steps):
                sf(); sb()
            s1.record(st)
            s1.synchronize()
            ts = torch.tensor([s0.elapsed_time(s1) / args.steps, err], dtype=torch.float64, device=dev)
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            out[f"{mode}_staged{n_stages}"] = {"step_ms": float(ts[0]), "edges_per_s": total / (float(ts[0]) * 1e-3),
                                               "max_abs_diff_vs_unstaged": float(ts[1]),
                                               "stage_rows_rank0": [int(p_.numel()) for p_ in sf.sp.recv_pos]}
            del sf, sb
        del f, b, ops_, new
        torch.cuda.empty_cache()
    if rank == 0:
        print("SCALE_COMPARE " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
