#!/usr/bin/env python
"""First GPU call of the next round: check and time the two single-GPU experiments that were written
without GPU access: the single-launch SpMM forward (`spmm_fused`) and the native
train step (`native_step.NativeTrainStep`).  Prints one JSON line per experiment.

    gpurun --timeout 300 -- 'python tools/round2_checks.py > gpurun_out/round2_checks.log 2>&1; tail -5 gpurun_out/round2_checks.log'
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gae_dgl_b200 as G  # noqa: E402
from gae_dgl_b200 import _lib, ops, synthetic  # noqa: E402
from gae_dgl_b200.graph import PackedGraphDataset, coo_to_csr_torch  # noqa: E402


def time_ms(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def check_fused_spmm(dev):
    out = {"experiment": "spmm_fused"}
    for scale, edges, d in ((16, 1_500_000, 64), (16, 1_500_000, 32), (22, 100_000_000, 64)):
        n = 1 << scale
        src, dst = synthetic.rmat_edges(scale, edges, seed=1, device=dev)
        rowptr, col = coo_to_csr_torch(src, dst, n)
        del src, dst
        plan = ops.build_hub_plan(rowptr, 512, bins=True)
        ops.order_segments_by_source(plan, rowptr, col)
        X = synthetic.hashed_normal(n, d, 2, device=dev)
        ws = plan.workspace(d, dev)
        Y0, Y1 = torch.empty_like(X), torch.full_like(X, float("nan"))
        _lib.set_tuning("spmm_fused", 0)
        ops.spmm(rowptr, col, X, plan, out=Y0, partial_ws=ws)
        t0 = time_ms(lambda: ops.spmm(rowptr, col, X, plan, out=Y0, partial_ws=ws), 10)
        _lib.set_tuning("spmm_fused", 1)
        ops.spmm(rowptr, col, X, plan, out=Y1, partial_ws=ws)
        t1 = time_ms(lambda: ops.spmm(rowptr, col, X, plan, out=Y1, partial_ws=ws), 10)
        _lib.set_tuning("spmm_fused", 0)
        out[f"scale{scale}_E{edges}_d{d}"] = {"separate_ms": t0, "fused_ms": t1, "bit_identical": bool(torch.equal(Y0, Y1)),
                                              "max_abs_diff": float((Y0 - Y1).abs().nan_to_num(nan=1e30).max())}
        del rowptr, col, X, Y0, Y1, ws, plan
        torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


def check_native_step(dev):
    from gae_dgl_b200.native_step import NativeTrainStep
    out = {"experiment": "native_step"}
    ds = synthetic.zinc_like_dataset(2048, seed=0)
    packed = PackedGraphDataset(ds, dev)
    rng = np.random.default_rng(0)
    batches = [rng.permutation(len(ds))[:256] for _ in range(210)]

    def make():
        torch.manual_seed(0)
        m = G.GAE(39, [32, 16]).to(dev)
        return m, torch.optim.Adam(m.parameters(), lr=1e-3, fused=True)

    # parity: same batches, same injected masks, autograd path vs native path
    m1, o1 = make()
    m2, o2 = make()
    native = NativeTrainStep(m2, o2)
    worst = 0.0
    for ids in batches[:6]:
        bg1, bg2 = packed.batch(ids), packed.batch(ids)
        mask = (torch.rand(bg1.number_of_nodes(), 16, device=dev) >= 0.1)
        l1 = m1.loss(bg1, mask=mask)
        o1.zero_grad(set_to_none=True)
        l1.backward()
        o1.step()
        l2 = native(bg2, mask=mask)
        worst = max(worst, abs(float(l1) - float(l2)) / abs(float(l1)))
    wdiff = max(float((a - b).abs().max()) for a, b in zip(m1.parameters(), m2.parameters()))
    native.sync_state()
    out["loss_rel_diff_max"] = worst
    out["weights_abs_diff_after_6_steps"] = wdiff
    out["optimizer_step_counts"] = [float(o1.state[p]["step"]) for p in m1.parameters()][:1] + \
                                   [float(o2.state[p]["step"]) for p in m2.parameters()][:1]

    def loop(step_fn):
        for ids in batches[:10]:
            step_fn(ids)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for ids in batches[10:]:
            last = step_fn(ids)
        float(last)
        return (time.perf_counter() - t0) * 1e3 / (len(batches) - 10)

    def autograd_step(ids):
        loss = m1.loss(packed.batch(ids))
        o1.zero_grad(set_to_none=True)
        loss.backward()
        o1.step()
        return loss

    out["autograd_ms_per_step"] = loop(autograd_step)
    out["native_ms_per_step"] = loop(lambda ids: native(packed.batch(ids)))
    print(json.dumps(out), flush=True)


def main():
    dev = torch.device("cuda:0")
    for fn in (check_fused_spmm, check_native_step):
        try:
            fn(dev)
        except Exception as exc:  # noqa: BLE001  -- report and go on to the next experiment
            import traceback
            traceback.print_exc()
            print(json.dumps({"experiment": fn.__name__, "error": repr(exc)}), flush=True)


if __name__ == "__main__":
    main()
