#!/usr/bin/env python
"""Time the fused decoder (loss + gradient) on a Pubmed-shaped / ZINC-batch-shaped embedding for
each combination of the decoder tuning knobs (CUDA events, L2-resident inputs: Z is 1.3 MB)."""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gae_dgl_b200 import _lib, ops, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--rows", type=str, default="1,2")
    ap.add_argument("--splits", type=str, default="0")
    ap.add_argument("--mma", type=str, default="0,1")
    ap.add_argument("--out", type=str, default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    results = []
    for name in ("pubmed", "cora"):
        g, _ = synthetic.planetoid_like(name, seed=0)
        g.to(dev)
        c, t = g.csr(), g.csr_t()
        n = g.number_of_nodes()
        Z = (0.3 * torch.randn(n, 16, generator=torch.Generator().manual_seed(0))).to(dev)
        ref = None
        for rows, splits, mma in itertools.product(*[[int(v) for v in s.split(",")]
                                                     for s in (args.rows, args.splits, args.mma)]):
            if mma and rows != 2:
                continue
            _lib.set_tuning("dec_mma", mma)
            _lib.set_tuning("dec_rows", rows)
            _lib.set_tuning("dec_splits", splits)
            for _ in range(3):
                loss, dz = ops.decoder_bce(Z, c.rowptr, c.col, t.rowptr, t.col, 100.0, True, True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.iters):
                loss, dz = ops.decoder_bce(Z, c.rowptr, c.col, t.rowptr, t.col, 100.0, True, True)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1) / args.iters
            if ref is None:
                ref = (float(loss), dz.clone())
            r = {"graph": name, "n": n, "dec_rows": rows, "dec_splits": splits, "dec_mma": mma, "ms": ms,
                 "pairs_per_s": n * n / (ms * 1e-3), "loss": float(loss), "loss_diff_vs_first": abs(float(loss) - ref[0]),
                 "grad_maxdiff_vs_first": float((dz - ref[1]).abs().max())}
            results.append(r)
            print(json.dumps(r), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
