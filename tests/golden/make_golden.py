"""Generates tests/golden/gae_small.npz from the CPU oracle (fp64 ground truth + fp32 like-for-like).

The reference itself cannot be imported here (dgl is not installed, no network), so these
vectors pin the ORACLE against regressions and give the GPU parity test a committed fixture;
they are not outputs of DGL.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gae_oracle as O  # noqa: E402


def main():
    g = torch.Generator().manual_seed(1234)
    n, e, f = 97, 400, 39
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n, (e,), generator=g)
    # duplicates, a self loop, an isolated node (n-1 has no in-edges), one hub row
    src = torch.cat([src, torch.tensor([3, 3, 5]), torch.randint(0, n, (60,), generator=g)])
    dst = torch.cat([dst, torch.tensor([7, 7, 5]), torch.full((60,), 11)])
    keep = dst != n - 1
    src, dst = src[keep], dst[keep]
    rowptr, col = O.coo_to_csr(src, dst, n)
    X = torch.randn(n, f, generator=g)
    torch.manual_seed(7)
    model = O.OracleGAE(f, [32, 16])
    weights = [(l.apply_mod.linear.weight.detach(), l.apply_mod.linear.bias.detach()) for l in model.layers]
    mask = (torch.rand(n, 16, generator=g) >= 0.1)
    out = {"src": src.numpy(), "dst": dst.numpy(), "n": n, "rowptr": rowptr.numpy(), "col": col.numpy(),
           "X": X.numpy(), "mask": mask.numpy()}
    for i, (W, b) in enumerate(weights):
        out[f"W{i}"] = W.numpy()
        out[f"b{i}"] = b.numpy()
    out["spmm64"] = O.spmm_sum(rowptr, col, X.double()).numpy()
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        loss, z, grads = O.train_step(rowptr, col, X, weights, mask, p=0.1, dtype=dt)
        out[f"loss{tag}"] = loss.numpy()
        out[f"z{tag}"] = z.numpy()
        for i, (gW, gb) in enumerate(grads):
            out[f"gW{i}_{tag}"] = gW.numpy()
            out[f"gb{i}_{tag}"] = gb.numpy()
    adj = O.dense_adj_from_csr(rowptr, col)
    out["pos_weight"] = O.pos_weight_inductive(adj).numpy()
    out["pos_weight_transductive"] = O.pos_weight_transductive(adj).numpy()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gae_small.npz"), **out)
    print("loss32", out["loss32"], "loss64", out["loss64"], "pw", out["pos_weight"])


if __name__ == "__main__":
    main()
