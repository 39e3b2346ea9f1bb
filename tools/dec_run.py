#!/usr/bin/env python
"""Runs the fused decoder at the Pubmed shape a few times (target of ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gae_dgl_b200 import _lib, ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
name = sys.argv[1] if len(sys.argv) > 1 else "pubmed"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for kv in filter(None, (sys.argv[3] if len(sys.argv) > 3 else "").split(",")):
    k, v = kv.split("=")
    _lib.set_tuning(k, int(v))
g, X = synthetic.planetoid_like(name, seed=0)
g.to(dev)
c, t = g.csr(), g.csr_t()
Zd = torch.randn(g.number_of_nodes(), 16, device=dev) * 0.3
for _ in range(iters):
    loss, dZ = ops.decoder_bce(Zd, c.rowptr, c.col, t.rowptr, t.col, 5.0, want_loss=True, want_grad=True)
torch.cuda.synchronize()
print(float(loss))
