#!/usr/bin/env python
"""Multi-GPU parity check (run under torch.distributed.run): partitioned SpMM fwd/bwd with both
exchange mechanisms vs the CPU oracle on the full graph."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from gae_dgl_b200 import parallel, synthetic
    from oracle import c_spmm
    from oracle import gae_oracle as O
    scale, n_edges, d = 16, 2_000_000, 64
    n = 1 << scale
    ok = True
    # (exchange, stages): "halo" = staged one-sided push with device flags, overlapped with `stages` row blocks
    # stages > 0: row blocks; stages = 0 / -1: popularity classes with thresholds (8,) / (32, 4);
    # stages = -2 / -3: hot rows copied + the tail folded by its owners, thresholds (16, 2) / (4, 3)
    for exchange, stages in (("halo", -2), ("halo", -3), ("halo", 0), ("halo", -1), ("halo", 1), ("halo", 4), ("halo", 8), ("nccl", 1), ("push", 1),
                             ("p2p", 1)):
        kw = {}
        if exchange == "halo":
            kw = (dict(kind="blocks") if stages > 0 else
                  dict(kind="classes", thresholds=(8,) if stages == 0 else (32, 4)) if stages >= -1 else
                  dict(kind="fold", thresholds=(16, 2) if stages == -2 else (4, 3)))
        part = parallel.build_rmat_partition(scale, n_edges, seed=1, d=d, device=dev, exchange=exchange,
                                             stages=max(stages, 1), **kw)
        Yf = part.fwd().clone()
        Yb = part.bwd().clone()
        if exchange == "halo":
            # epochs 2..4: the flags must order successive calls as well (write-after-read on the halo region);
            # the local rows are changed and restored in between so that a stale halo row would show
            for k in range(3):
                part.fwd_op.X_local.mul_(2.0 if k == 0 else 1.0)
                part.fwd()
                part.bwd()
            part.fwd_op.X_local.mul_(0.5)
            Yf2 = part.fwd().clone()
            torch.cuda.synchronize()
            part.fwd_op.check()
            part.bwd_op.check()
            assert torch.equal(Yf, Yf2), "halo exchange: a later epoch differs from the first"
        torch.cuda.synchronize()
        bounds = parallel.block_bounds(n, world)
        lo, hi = bounds[rank], bounds[rank + 1]
        S, D = synthetic.rmat_edges(scale, n_edges, seed=1)
        rp, col = O.coo_to_csr(S, D, n)
        rpt, colt = O.coo_to_csr(D, S, n)
        X = synthetic.hashed_normal(n, d, 2).numpy()
        dY = synthetic.hashed_normal(n, d, 3).numpy()
        ref_f = torch.from_numpy(c_spmm.spmm_f64acc(rp.numpy(), col.numpy(), X))[lo:hi]
        ref_b = torch.from_numpy(c_spmm.spmm_f64acc(rpt.numpy(), colt.numpy(), dY))[lo:hi]
        ef = float((Yf.double().cpu() - ref_f).abs().max() / max(float(ref_f.abs().max()), 1.0))
        eb = float((Yb.double().cpu() - ref_b).abs().max() / max(float(ref_b.abs().max()), 1.0))
        print(f"[rank {rank}] exchange={exchange} stages={stages} fwd_err={ef:.2e} bwd_err={eb:.2e} halo={part.halo_rows} "
              f"local_rows={part.local_rows} local_edges={part.local_edges}", flush=True)
        ok = ok and ef < 1e-5 and eb < 1e-5
        del part
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if int(t) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
