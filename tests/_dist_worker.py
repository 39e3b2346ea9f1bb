"""Worker for the world_size-2 gloo tests (spawned; must be importable)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _cpu_pack(X, idx, out):
    out.copy_(X.index_select(0, idx))
    return out


def _cpu_spmm(rowptr, col, X, plan, out, ws):
    # test double for the CUDA kernel: the oracle is the checker of the exchange logic here
    from oracle import gae_oracle as O
    out.copy_(O.spmm_sum(rowptr, col, X))
    return out


def run(rank, world, port, scale, n_edges, d, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gae_dgl_b200 import parallel, synthetic
        from oracle import gae_oracle as O
        n = 1 << scale
        bounds = parallel.block_bounds(n, world)
        per = (n_edges + world - 1) // world
        first = rank * per
        count = max(0, min(per, n_edges - first))
        src, dst = synthetic.rmat_edges(scale, count, seed=1, first_edge=first)
        fs, fd = parallel.route_edges(src, dst, dst, bounds)
        assert int(fd.min()) >= bounds[rank] and int(fd.max()) < bounds[rank + 1]
        hp = parallel.build_halo_plan(fs, fd, n, rank, world)
        bs, bd = parallel.route_edges(dst, src, src, bounds)
        hpb = parallel.build_halo_plan(bs, bd, n, rank, world)
        # every edge arrived exactly once
        tot = torch.tensor([hp.n_edges, hpb.n_edges])
        dist.all_reduce(tot)
        assert tot.tolist() == [n_edges, n_edges], tot
        lo, hi = bounds[rank], bounds[rank + 1]
        Xg = synthetic.hashed_normal(n, d, 2)                     # every rank can rebuild the global matrix
        out = {}
        for name, plan, (gs, gd) in (("fwd", hp, ("src", "dst")), ("bwd", hpb, ("dst", "src"))):
            op = parallel.PartitionedSpMM(plan, d, "nccl", pack_fn=_cpu_pack, spmm_fn=_cpu_spmm)
            op.X_local.copy_(Xg[lo:hi])
            Y = op().clone()
            # halo rows hold exactly the owners' rows
            assert torch.equal(op.X_halo, Xg[plan.halo_ids])
            assert sum(plan.recv_counts) == plan.n_halo and int(plan.col.max()) < plan.n_local + plan.n_halo
            out[name] = Y
            # staged exchange pipelined with row-block SpMM (parallel_staged): same result, and block b
            # reads only halo rows that stages <= b have delivered (the halo is poisoned beforehand)
            from gae_dgl_b200 import parallel_staged as PS
            for n_stages in (1, 2, 4, 7):
                st = PS.StagedPartitionedSpMM(op, n_stages, overlap=False)
                sp = st.sp
                assert sp.row_bounds[0] == 0 and sp.row_bounds[-1] == plan.n_local
                assert sorted(torch.cat(sp.recv_pos).tolist()) == list(range(plan.n_halo))      # every halo row once
                assert sum(int(i.numel()) for i in sp.send_idx) == int(plan.send_idx.numel())
                edges_per_block = [int(r[-1]) for r in sp.sub_rowptr]
                assert sum(edges_per_block) == plan.n_edges
                op.X_halo.fill_(float("nan"))
                op.Y.fill_(float("nan"))
                Ys = st()
                assert torch.equal(Ys, Y), (name, n_stages)
                assert torch.equal(op.X_halo, Xg[plan.halo_ids])
                # first-use tags: a halo row's stage is the block of the first local row that references it
                deg = plan.rowptr[1:] - plan.rowptr[:-1]
                erow = torch.repeat_interleave(torch.arange(plan.n_local), deg)
                col64 = plan.col.to(torch.int64)
                for h in range(0, plan.n_halo, max(1, plan.n_halo // 50)):
                    first_row = int(erow[col64 == plan.n_local + h].min())
                    blk = max(b for b in range(sp.n_stages) if sp.row_bounds[b] <= first_row)
                    assert int(sp.halo_stage[h]) == blk
        # global ground truth from the full edge stream
        S, D = synthetic.rmat_edges(scale, n_edges, seed=1)
        rp, col = O.coo_to_csr(S, D, n)
        ref_f = O.spmm_sum(rp, col, Xg.double())[lo:hi]
        rpt, colt = O.coo_to_csr(D, S, n)
        ref_b = O.spmm_sum(rpt, colt, Xg.double())[lo:hi]
        ef = float((out["fwd"].double() - ref_f).abs().max() / max(float(ref_f.abs().max()), 1.0))
        eb = float((out["bwd"].double() - ref_b).abs().max() / max(float(ref_b.abs().max()), 1.0))
        assert ef < 1e-5 and eb < 1e-5, (ef, eb)
        with open(os.path.join(result_dir, f"ok_{rank}"), "w") as f:
            f.write(f"{ef} {eb} halo={hp.n_halo} local={hp.n_local}")
    finally:
        dist.destroy_process_group()
