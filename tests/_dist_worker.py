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
            # staged one-sided exchange (HaloSpMM's plan, built by the gae_halo_*_host helpers): the push
            # lists are executed here by a gloo emulation of the push kernel (entries grouped by peer,
            # rows + destination indices through all-to-all-v), stage by stage, followed by the row-block
            # SpMM of that stage.  Same result as the unstaged op, and block b reads only halo rows that
            # stages <= b have delivered (the halo is poisoned beforehand).
            plans = [parallel.build_stage_plan(plan, k) for k in (1, 2, 4, 7)]
            plans += [parallel.build_class_plan(plan, t) for t in ((8,), (16, 2), (10 ** 9,))]
            for sp in plans:
                n_stages = sp.n_stages
                assert int(sp.stage_ptr[0]) == 0 and int(sp.stage_ptr[-1]) == int(plan.send_idx.numel())
                assert sorted(sp.push_src.tolist()) == sorted(plan.send_idx.tolist())
                assert sum(int(r[-1]) for r in sp.sub_rowptr) == plan.n_edges
                assert sorted(sp.halo_pos.tolist()) == list(range(plan.n_local, plan.n_local + plan.n_halo))
                if sp.kind == "blocks":
                    assert sp.row_bounds[0] == 0 and sp.row_bounds[-1] == plan.n_local
                else:
                    # class-major halo layout; a piece reads local columns (piece 0 only) and its own class
                    for k in range(sp.n_stages):
                        c = sp.sub_col[k].to(torch.int64)
                        hcols = c[c >= plan.n_local]
                        assert k == 0 or hcols.numel() == c.numel()
                        inv = torch.empty(plan.n_halo, dtype=torch.int64)
                        inv[sp.halo_pos - plan.n_local] = torch.arange(plan.n_halo)
                        assert bool((sp.halo_stage[inv[hcols - plan.n_local]] == k).all())
                ext = torch.full((plan.n_local + plan.n_halo, d), float("nan"))
                ext[:plan.n_local] = op.X_local
                Ys = torch.full_like(op.Y, float("nan"))
                landed = torch.zeros(plan.n_local + plan.n_halo, dtype=torch.int32)
                for st in range(sp.n_stages):
                    e0, e1 = int(sp.stage_ptr[st]), int(sp.stage_ptr[st + 1])
                    peer = sp.push_peer[e0:e1].to(torch.int64)
                    order = torch.argsort(peer, stable=True)
                    cnt = torch.bincount(peer, minlength=world)
                    rcnt = torch.empty_like(cnt)
                    dist.all_to_all_single(rcnt, cnt)
                    rows = op.X_local.index_select(0, sp.push_src[e0:e1][order])
                    dsts = sp.push_dst[e0:e1][order].contiguous()
                    got_rows = torch.empty((int(rcnt.sum()), d))
                    got_dst = torch.empty(int(rcnt.sum()), dtype=torch.int64)
                    parallel.all_to_all_v(got_rows, rows, rcnt.tolist(), cnt.tolist())
                    parallel.all_to_all_v(got_dst, dsts, rcnt.tolist(), cnt.tolist())
                    assert got_dst.numel() == 0 or (int(got_dst.min()) >= plan.n_local and
                                                    int(got_dst.max()) < plan.n_local + plan.n_halo)
                    ext[got_dst] = got_rows
                    landed[got_dst] += 1
                    r0, nr = sp.piece_rows[st]
                    if nr > 0:
                        part_y = torch.empty((nr, d))
                        _cpu_spmm(sp.sub_rowptr[st], sp.sub_col[st], ext, None, part_y, None)
                        assert not torch.isnan(part_y).any(), (name, sp.kind, n_stages, st)   # only delivered rows were read
                        if sp.accumulate[st]:
                            Ys[r0:r0 + nr] += part_y
                        else:
                            Ys[r0:r0 + nr] = part_y
                assert torch.equal(landed[plan.n_local:], torch.ones(plan.n_halo, dtype=torch.int32))   # every halo row once
                tol = 0.0 if sp.kind == "blocks" else 1e-5 * max(float(Y.abs().max()), 1.0)   # classes re-associate the sum
                assert float((Ys - Y).abs().max()) <= tol, (name, sp.kind, n_stages)
                assert torch.equal(ext[sp.halo_pos], Xg[plan.halo_ids])
                if sp.kind == "blocks":
                    # first-use tags: a halo row's stage is the block of the first local row that references it
                    deg = plan.rowptr[1:] - plan.rowptr[:-1]
                    erow = torch.repeat_interleave(torch.arange(plan.n_local), deg)
                    col64 = plan.col.to(torch.int64)
                    for h in range(0, plan.n_halo, max(1, plan.n_halo // 50)):
                        first_row = int(erow[col64 == plan.n_local + h].min())
                        blk = max(b for b in range(sp.n_stages) if sp.row_bounds[b] <= first_row)
                        assert int(sp.halo_stage[h]) == blk
                else:
                    cnt_ref = torch.bincount(plan.col.to(torch.int64)[plan.col >= plan.n_local] - plan.n_local,
                                             minlength=plan.n_halo)
                    assert bool((sp.halo_stage[cnt_ref >= 10 ** 9] == 0).all())
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
