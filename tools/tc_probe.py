#!/usr/bin/env python
"""Check of the tcgen05 / TMEM decoder pass on a B200 (one GPU): per-tile products against fp64 host products
(gae_decoder_tile_probe_f32), then the whole fused decoder against the mma.sync path and the fp64 closed form,
then timings at the Pubmed shape.  Dumps the raw tiles to gpurun_out/tc_probe.npz for offline inspection."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from gae_dgl_b200 import _lib, ops, synthetic  # noqa: E402
import gae_dgl_b200 as G  # noqa: E402


def probe(Zd, ti, tj):
    lib = _lib.load()
    n, d = Zd.shape
    S = torch.full((128, 128), float("nan"), device=Zd.device)
    Gi = torch.full((128, 16), float("nan"), device=Zd.device)
    Gj = torch.full((128, 16), float("nan"), device=Zd.device)
    t = ctypes.c_int32(-1)
    rc = lib.gae_decoder_tile_probe_f32(ops._ptr(Zd), Zd.stride(0), n, d, ti, tj, ops._ptr(S), ops._ptr(Gi), ops._ptr(Gj),
                                        ctypes.byref(t), ops._stream())
    _lib.check(rc, "gae_decoder_tile_probe_f32")
    return S.cpu(), Gi.cpu(), Gj.cpu(), t.value


def rel(a, b):
    return float((a.double() - b).abs().max() / max(float(b.abs().max()), 1e-30))


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    # which tensor-core kernel: 1 = TF32 form (decoder_tc.cu), 2 = fp16-split pipelined form (decoder_tc16.cu)
    TC = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    _lib.set_tuning("dec_tc", TC)
    out, dump = {}, {}
    ok = True
    for n, d, tiles in ((1000, 16, [(0, 0), (0, 1), (2, 7), (7, 7)]), (700, 13, [(0, 3), (5, 5)])):
        g = torch.Generator().manual_seed(n)
        Zd = ops.alloc_rows(n, d, dev)
        Zd.copy_((torch.randn(n, d, generator=g) * 0.7).to(dev))
        Z = torch.zeros(1024, 16, dtype=torch.float64)
        Z[:n, :d] = Zd.double().cpu()
        for ti, tj in tiles:
            S, Gi, Gj, to = probe(Zd, ti, tj)
            ZI, ZJ = Z[128 * ti:128 * ti + 128], Z[128 * tj:128 * tj + 128]
            S_ref = ZI @ ZJ.t()
            sig = torch.sigmoid(S_ref)
            rows_ok = (torch.arange(128) + 128 * ti < n).double()[:, None]
            keys_ok = (torch.arange(128) + 128 * tj < n).double()[None, :]
            sig = sig * rows_ok * keys_ok
            e = {"timeouts": to, "S": rel(S, S_ref), "G_i": rel(Gi, sig @ ZJ)}
            if ti != tj:
                e["G_j"] = rel(Gj, sig.t() @ ZI)
            out[f"n{n}_d{d}_tile{ti}_{tj}"] = e
            ok = ok and to == 0 and all(v < 2e-5 for k, v in e.items() if k != "timeouts")
            dump[f"n{n}_t{ti}_{tj}_S"] = S.numpy()
            dump[f"n{n}_t{ti}_{tj}_Gi"] = Gi.numpy()
            dump[f"n{n}_t{ti}_{tj}_Gj"] = Gj.numpy()
        dump[f"n{n}_Z"] = Z.numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "tc_probe.npz"), **dump)
    print(json.dumps({"probe": out, "probe_ok": ok}), flush=True)

    # whole decoder: tcgen05 path vs mma.sync path vs fp64 closed form
    from oracle import gae_oracle as O
    full = {}
    for n, d, e, zmul in ((700, 16, 3000, 0.5), (1500, 16, 6000, 0.5), (2708, 16, 10556, 0.5), (3000, 7, 9000, 0.5),
                          (900, 16, 4000, 3.0e4), (900, 16, 4000, 1.0e-3)):
        gen = torch.Generator().manual_seed(n + d)
        src = torch.randint(0, n, (e,), generator=gen)
        dst = torch.randint(0, n, (e,), generator=gen)
        g = G.DGLGraph((src.numpy(), dst.numpy(), n))
        g.to(dev)
        c, t = g.csr(), g.csr_t()
        Zd = (torch.randn(n, d, generator=gen) * zmul).to(dev)
        res = {}
        for knob in (1, 0):
            _lib.set_tuning("dec_tc", TC if knob else 0)
            loss, dZ = ops.decoder_bce(Zd, c.rowptr, c.col, t.rowptr, t.col, 7.5, want_loss=True, want_grad=True)
            res[knob] = (float(loss), dZ.double().cpu())
        _lib.set_tuning("dec_tc", TC)
        zd64 = Zd.double().cpu().requires_grad_(True)
        ref = O.bce_loss_sparse_form(zd64, c.rowptr.cpu(), c.col.cpu(), 7.5)
        ref.backward()
        gref = zd64.grad
        full[f"n{n}_d{d}_z{zmul:g}"] = {
            "loss_tc": res[1][0], "loss_mma": res[0][0], "loss_ref": float(ref),
            "loss_rel_tc": abs(res[1][0] - float(ref)) / abs(float(ref)),
            "grad_rel_tc": float((res[1][1] - gref).abs().max() / gref.abs().max()),
            "grad_rel_mma": float((res[0][1] - gref).abs().max() / gref.abs().max())}
        ok = ok and full[f"n{n}_d{d}_z{zmul:g}"]["loss_rel_tc"] < 1e-5 and full[f"n{n}_d{d}_z{zmul:g}"]["grad_rel_tc"] < 1e-5
    print(json.dumps({"decoder": full, "all_ok": ok}), flush=True)

    # timing at the Pubmed shape
    g, X = synthetic.planetoid_like("pubmed", seed=0)
    g.to(dev)
    c, t = g.csr(), g.csr_t()
    Zd = torch.randn(19717, 16, device=dev) * 0.3
    tim = {}
    for knob in (2, 1, 0):
        _lib.set_tuning("dec_tc", knob)
        for _ in range(3):
            ops.decoder_bce(Zd, c.rowptr, c.col, t.rowptr, t.col, 5.0, want_loss=True, want_grad=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            loss, _ = ops.decoder_bce(Zd, c.rowptr, c.col, t.rowptr, t.col, 5.0, want_loss=True, want_grad=True)
        e1.record()
        e1.synchronize()
        tim[{2: "tc16", 1: "tc_tf32", 0: "mma_sync"}[knob]] = {"ms": e0.elapsed_time(e1) / 20, "loss": float(loss)}
    _lib.set_tuning("dec_tc", TC)
    print(json.dumps({"pubmed_decoder_ms": tim}), flush=True)


if __name__ == "__main__":
    main()
