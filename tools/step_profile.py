#!/usr/bin/env python
"""Time the small-graph train steps (Pubmed-shaped transductive, ZINC-shaped batch) and, under
ncu --metrics gpu__time_duration.sum, list every kernel of one step."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import gae_dgl_b200 as G  # noqa: E402
from gae_dgl_b200 import synthetic  # noqa: E402


def run(name, g, X, in_dim, lr, steps, transductive, graphed=True):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = G.GAE(in_dim, [32, 16]).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=lr, fused=True)
    g.to(dev)
    Xd = X.to(dev)
    pw = G.pos_weight_of(g, transductive=transductive)

    def step():
        g.ndata["h"] = Xd
        loss = model.loss(g, pos_weight=pw)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push(name)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    e1.synchronize()
    torch.cuda.nvtx.range_pop()
    wall = (time.perf_counter() - t0) / steps * 1e3
    print(f"{name}: N={g.number_of_nodes()} E={g.number_of_edges()} {e0.elapsed_time(e1) / steps:.3f} ms/step (gpu events) "
          f"{wall:.3f} ms/step (wall)", flush=True)
    if graphed:
        from gae_dgl_b200.graphed import GraphedTrainStep
        opt2 = torch.optim.Adam(model.parameters(), lr=lr, capturable=True, fused=True)

        def loss_fn():
            g.ndata["h"] = Xd
            return model.loss(g, pos_weight=pw)

        gs = GraphedTrainStep(model, opt2, loss_fn)
        for _ in range(3):
            gs()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            gs()
        e1.record()
        e1.synchronize()
        print(f"{name}: graphed {e0.elapsed_time(e1) / steps:.3f} ms/step  loss={float(gs.loss):.5f}", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--which", type=str, default="pubmed,zinc,cora")
    ap.add_argument("--tune", type=str, default="")
    args = ap.parse_args()
    from gae_dgl_b200 import _lib
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        _lib.set_tuning(k, int(v))
    if "pubmed" in args.which:
        g, X = synthetic.planetoid_like("pubmed", seed=0)
        run("pubmed", g, X, 500, 1e-2, args.steps, True)
    if "cora" in args.which:
        g, X = synthetic.planetoid_like("cora", seed=0)
        run("cora", g, X, 1433, 1e-2, args.steps, True)
    if "zinc" in args.which:
        ds = synthetic.zinc_like_dataset(256, seed=0)
        bg = G.batch(ds, device="cuda:0")
        X = bg.ndata["h"].clone()
        run("zinc_b256", bg, X, 39, 1e-3, args.steps, False)


if __name__ == "__main__":
    main()
