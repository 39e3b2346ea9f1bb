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
            # folded exchange (kind "fold", the default of HaloSpMM): hot rows copied, the tail summed by its
            # owners.  Emulated the same way: the sums I owe (pre CSR over my local rows) go into staging rows
            # behind my halo region, then the two stages are pushed and the two pieces aggregated.
            for hot, fold in ((4, 2), (16, 2), (2, 3), (1, 2), (10 ** 9, 2), (10 ** 9, 10 ** 9)):
                sp = parallel.build_fold_plan(plan, hot, fold)
                assert sp.kind == "fold" and sp.n_stages == 2 and sp.accumulate == [False, True] and sp.pre_stage == 1
                n_ext, n_pre, base = sp.n_ext, sp.n_pre, plan.n_local + sp.n_ext
                st_ = sp.stats
                assert st_["rows_in"] == n_ext == st_["hot_rows"] + st_["cold_rows"] + st_["folded_rows_in"]
                # every one of my edges is aggregated exactly once: by me (pieces) or by an owner (folded)
                assert int(sp.sub_rowptr[0][-1]) + st_["cold_edges"] + st_["folded_edges_in"] == plan.n_edges
                assert int(sp.sub_rowptr[1][-1]) == st_["cold_edges"] + st_["folded_rows_in"]
                tot = torch.tensor([n_ext, int(sp.stage_ptr[-1]), st_["folded_edges_in"], st_["folded_edges_out"]])
                dist.all_reduce(tot)
                assert int(tot[0]) == int(tot[1]) and int(tot[2]) == int(tot[3])
                assert (hot, fold) != (4, 2) or int(tot[2]) > 0          # the small test graph does fold something
                if hot == 1:
                    assert n_ext == plan.n_halo and n_pre == 0 and st_["cold_rows"] == 0
                if fold >= 10 ** 9:
                    assert n_ext == plan.n_halo and n_pre == 0 and st_["folded_rows_in"] == 0
                ext = torch.full((base + n_pre, d), float("nan"))
                ext[:plan.n_local] = op.X_local
                if n_pre:
                    assert int(sp.pre_col.max()) < plan.n_local and int(sp.pre_rowptr[-1]) == sp.pre_col.numel()
                    pre = torch.empty((n_pre, d))
                    _cpu_spmm(sp.pre_rowptr, sp.pre_col, ext[:plan.n_local], None, pre, None)
                    ext[base:] = pre
                Ys = torch.full_like(op.Y, float("nan"))
                landed = torch.zeros(base, dtype=torch.int32)
                for st in range(2):
                    e0, e1 = int(sp.stage_ptr[st]), int(sp.stage_ptr[st + 1])
                    srcs = sp.push_src[e0:e1]
                    assert st >= sp.pre_stage or srcs.numel() == 0 or int(srcs.max()) < plan.n_local
                    assert srcs.numel() == 0 or bool(((srcs < plan.n_local) | (srcs >= base)).all())
                    peer = sp.push_peer[e0:e1].to(torch.int64)
                    order = torch.argsort(peer, stable=True)
                    cnt = torch.bincount(peer, minlength=world)
                    rcnt = torch.empty_like(cnt)
                    dist.all_to_all_single(rcnt, cnt)
                    rows = ext.index_select(0, srcs[order])
                    dsts = sp.push_dst[e0:e1][order].contiguous()
                    got_rows = torch.empty((int(rcnt.sum()), d))
                    got_dst = torch.empty(int(rcnt.sum()), dtype=torch.int64)
                    parallel.all_to_all_v(got_rows, rows, rcnt.tolist(), cnt.tolist())
                    parallel.all_to_all_v(got_dst, dsts, rcnt.tolist(), cnt.tolist())
                    assert got_dst.numel() == 0 or (int(got_dst.min()) >= plan.n_local and int(got_dst.max()) < base)
                    ext[got_dst] = got_rows
                    landed[got_dst] += 1
                    part_y = torch.empty((plan.n_local, d))
                    _cpu_spmm(sp.sub_rowptr[st], sp.sub_col[st], ext[:base], None, part_y, None)
                    assert not torch.isnan(part_y).any(), (name, "fold", hot, fold, st)
                    Ys = part_y if st == 0 else Ys + part_y
                assert torch.equal(landed[plan.n_local:], torch.ones(n_ext, dtype=torch.int32))
                tol = 1e-5 * max(float(Y.abs().max()), 1.0)
                assert float((Ys - Y).abs().max()) <= tol, (name, "fold", hot, fold)
                # the copied rows are the owners' rows; rows that arrive folded have no slot
                kept = sp.halo_pos >= 0
                assert torch.equal(ext[sp.halo_pos[kept]], Xg[plan.halo_ids[kept]])
                assert int(kept.sum()) == st_["hot_rows"] + st_["cold_rows"]
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
