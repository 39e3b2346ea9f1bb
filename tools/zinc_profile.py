#!/usr/bin/env python
"""Host-side profile of the inductive (ZINC-shaped) train step: where the 0.6 ms per step go.
cProfile over the same loop bench.py's zinc leg times (device collation + fused step + Adam)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import gae_dgl_b200 as G  # noqa: E402
from gae_dgl_b200 import synthetic  # noqa: E402
from gae_dgl_b200.graph import PackedGraphDataset  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    steps, warmup, bs = 200, 10, 256
    ds = synthetic.zinc_like_dataset(4096, seed=0)
    packed = PackedGraphDataset(ds, dev)
    torch.manual_seed(0)
    model = G.GAE(39, [32, 16]).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    rng = np.random.default_rng(0)
    batches = [rng.permutation(len(ds))[:bs] for _ in range(steps + warmup)]

    def step(ids):
        loss = model.loss(packed.batch(ids))
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for ids in batches[:warmup]:
        step(ids)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for ids in batches[warmup:]:
        last = step(ids)
    float(last)
    print(f"plain loop: {(time.perf_counter() - t0) * 1e3 / steps:.3f} ms/step")
    # phases, each synchronised (upper bounds: they serialise host and device)
    for name, fn in (("batch", lambda ids: packed.batch(ids)),):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for ids in batches[warmup:]:
            fn(ids)
        torch.cuda.synchronize()
        print(f"{name} only: {(time.perf_counter() - t0) * 1e3 / steps:.3f} ms/step")
    bgs = [packed.batch(ids) for ids in batches[warmup:warmup + 50]]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for bg in bgs:
        loss = model.loss(bg)
    torch.cuda.synchronize()
    print(f"loss fwd only: {(time.perf_counter() - t0) * 1e3 / 50:.3f} ms/step")
    t0 = time.perf_counter()
    for bg in bgs:
        bg.ndata["h"] = packed.feat_all[:bg.number_of_nodes()]
        loss = model.loss(bg)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    print(f"loss+bwd+adam (no collation): {(time.perf_counter() - t0) * 1e3 / 50:.3f} ms/step")
    pr = cProfile.Profile()
    pr.enable()
    for ids in batches[warmup:]:
        last = step(ids)
    float(last)
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(28)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
