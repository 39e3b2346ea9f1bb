#!/usr/bin/env python
"""ZINC-shaped inductive loop: host collation vs device-resident packed collation (ms per batch)."""
import os
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
    ds = synthetic.zinc_like_dataset(4096, seed=0)
    packed = PackedGraphDataset(ds, dev)
    torch.manual_seed(0)
    model = G.GAE(39, [32, 16]).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    rng = np.random.default_rng(0)
    batches = [rng.permutation(4096)[:256] for _ in range(30)]

    def train(bg):
        loss = model.loss(bg)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for name, collate in (("host", lambda ids: G.batch([ds[i] for i in ids], device=dev)), ("packed", packed.batch)):
        for ids in batches[:3]:
            train(collate(ids))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for ids in batches:
            collate(ids)
        torch.cuda.synchronize()
        tc = (time.perf_counter() - t0) / len(batches) * 1e3
        t0 = time.perf_counter()
        for ids in batches:
            l = train(collate(ids))
        l.item()
        tt = (time.perf_counter() - t0) / len(batches) * 1e3
        edges = np.mean([packed.edges[i].sum() for i in batches])
        print(f"{name}: collate {tc:.3f} ms/batch, collate+train step {tt:.3f} ms/batch, {edges / tt * 1e3:.3e} edges/s "
              f"(batch=256, mean {edges:.0f} edges)", flush=True)


if __name__ == "__main__":
    main()
