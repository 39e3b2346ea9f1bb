#!/usr/bin/env python
"""Times the fused decoder (loss + gradient) at a Planetoid-like shape for a list of tuning settings.
usage: dec_time.py pubmed "dec_tc=2" "dec_tc=2,dec_splits=6" ..."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from gae_dgl_b200 import _lib, ops, synthetic  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
name = sys.argv[1]
if name.startswith("n") and name[1:].isdigit():          # "n6000": a random graph with that many vertices, 5 edges per vertex
    import gae_dgl_b200 as G
    nn = int(name[1:])
    gen = torch.Generator().manual_seed(nn)
    g = G.DGLGraph((torch.randint(0, nn, (5 * nn,), generator=gen).numpy(), torch.randint(0, nn, (5 * nn,), generator=gen).numpy(), nn))
else:
    g, X = synthetic.planetoid_like(name, seed=0)
g.to(dev)
c, t = g.csr(), g.csr_t()
Zd = torch.randn(g.number_of_nodes(), 16, device=dev) * 0.3
base = {k: _lib.get_tuning(k) for k in ("dec_tc", "dec_splits", "dec_mma")}
for spec in sys.argv[2:]:
    for k, v in base.items():
        _lib.set_tuning(k, v)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        _lib.set_tuning(k, int(v))
    for _ in range(3):
        ops.decoder_bce(Zd, c.rowptr, c.col, t.rowptr, t.col, 5.0, want_loss=True, want_grad=True)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            loss, dZ = ops.decoder_bce(Zd, c.rowptr, c.col, t.rowptr, t.col, 5.0, want_loss=True, want_grad=True)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    print(json.dumps({"shape": name, "tuning": spec, "decoder_ms": best, "loss": float(loss), "dZ_absmax": float(dZ.abs().max())}), flush=True)
