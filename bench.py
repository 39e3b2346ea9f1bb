#!/usr/bin/env python
"""Benchmark of the GAE hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3] at N=1, configs[4] at N=8; weak scaling in between):
Graph500 R-MAT (a,b,c,d = .57,.19,.19,.05), scale 22+log2(N), |E| = 1e8 * N directed edges,
no dedup, vertex labels scrambled (Graph500), d = 64 fp32 features.  One "step" is the
encoder aggregation train step on that graph: SpMM forward Y = A X followed by SpMM backward
dX = A^T dY (the N^2 decoder is infeasible at |V| >= 4M and is not run).
metric = |E| / step time.  The Pubmed (configs[1]) and ZINC batch=256 (configs[2]) train steps
are reported in the same line under "pubmed" / "zinc".

--impl reference times the CPU oracle (the reference's DGL path cannot be installed here:
no dgl wheel, no network) with all host threads on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

D_FEAT = 64
BASE_SCALE = 22
BASE_EDGES = 100_000_000
METRIC = "GAE train-step edges/sec (RMAT SpMM fwd+bwd, d=64)"


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=30)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    p.add_argument("--base-scale", type=int, default=BASE_SCALE, help="log2 |V| per GPU")
    p.add_argument("--base-edges", type=int, default=BASE_EDGES, help="|E| per GPU")
    p.add_argument("--strong", action="store_true",
                   help="strong scaling: the FIXED configs[4] graph (scale 25, 8e8 edges) on N GPUs instead of 1e8 edges per GPU")
    p.add_argument("--cpu-seconds", type=float, default=45.0, help="time budget of a CPU leg's timed steps (full workload)")
    p.add_argument("--no-pubmed", action="store_true")
    p.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--exchange", type=str, default="auto", choices=["auto", "halo", "nccl", "p2p", "push"])
    p.add_argument("--tune", type=str, default="", help="comma list key=value for gae_set_tuning")
    p.add_argument("--stages", type=int, default=8,
                   help="N > 1, exchange 'halo': stages of the one-sided push overlapped with that many row-block SpMMs")
    p.add_argument("--push-ctas", type=int, default=0, help="N > 1, exchange 'halo': CTAs of the push kernel (0 = 64)")
    p.add_argument("--halo-kind", type=str, default="fold", choices=["fold", "classes", "blocks"],
                   help="exchange 'halo': hot rows copied + the tail folded by its owners (default), popularity classes "
                        "of copied rows, or row blocks")
    p.add_argument("--hot", type=str, default="", help="halo-kind fold: 'hot,fold' thresholds (default 16,2); classes: "
                                                       "descending reference-count thresholds, e.g. 32,4 (default 8)")
    return p.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def spmm_algorithmic_bytes(n_edges: int, n_rows: int, d: int) -> int:
    """SURVEY.md 8d primary (gather) model: per edge 4 B col + 4d B source row; per row 4d B
    output + 4 B rowptr  ->  260 B/edge + 260 B/row at d = 64."""
    return n_edges * (4 + 4 * d) + n_rows * (4 * d + 4)


# ------------------------------------------------------------------------------------------
# clocks sampling (pynvml, falling back to nvidia-smi)
# ------------------------------------------------------------------------------------------

class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake"}

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU legs (oracle)
# ------------------------------------------------------------------------------------------

def workload_of(args, n_gpus: int):
    """(scale, total_edges, scaling): BASELINE.json configs[3] at N = 1 and configs[4] at N = 8 with 1e8 edges per GPU
    in between (weak), or the fixed configs[4] graph at every N (--strong)."""
    if args.strong:
        return 25, 800_000_000, "strong"
    return args.base_scale + int(round(math.log2(n_gpus))), args.base_edges * n_gpus, "weak"


def config_of(scale: int, total_edges: int):
    """The workload description shared by both arms (identical dicts => the driver's same_config check)."""
    n = 1 << scale
    return {"workload": f"rmat_scale{scale}_E{total_edges}_d{D_FEAT}", "step": "spmm_fwd+spmm_bwd",
            "rmat": "a=.57 b=.19 c=.19 d=.05, no dedup, labels scrambled, seed 1",
            "l2": f"inputs {4 * D_FEAT * n / 1e6:.0f} MB features + {4 * total_edges / 1e6:.0f} MB indices per SpMM "
                  ">> 126 MB L2, no flush between iterations"}


def host_csr_pair(scale: int, n_edges: int):
    """CSR and CSR^T of the whole workload as host numpy arrays.  Set-up only (untimed): the edge stream is
    drawn and sorted with torch on the GPU when one is present (same bits as on the CPU, synthetic.py), which
    keeps the full-size CPU leg inside a few minutes; nothing of this package's kernels is involved."""
    from gae_dgl_b200 import synthetic
    from gae_dgl_b200.graph import coo_to_csr_torch
    n = 1 << scale
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    src, dst = synthetic.rmat_edges(scale, n_edges, seed=1, device=dev)
    out = []
    for a, b in ((src, dst), (dst, src)):
        rowptr, col = coo_to_csr_torch(a, b, n)
        out += [rowptr.cpu().numpy(), col.cpu().numpy()]
        del rowptr, col
    del src, dst
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    return out


def cpu_rmat_leg(scale: int, n_edges: int, budget_s: float, max_steps: int = 10):
    """Times the oracle's SpMM fwd + bwd (the reference's update_all(copy_src,sum), gae.py:28, and its
    adjoint, train_inductive.py:51) on the WHOLE workload -- every edge, every vertex -- with all host
    threads.  Two restatements are timed, torch.sparse CSR (MKL) and the OpenMP C loop; the faster is
    reported.  CPU steps are bounded by time, not by shrinking the graph."""
    from oracle import c_spmm
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n = 1 << scale
    rp, cl, rpt, clt = host_csr_pair(scale, n_edges)
    # feature VALUES do not affect timing; torch.rand is used on the host to keep set-up short
    X = torch.rand(n, D_FEAT, generator=torch.Generator().manual_seed(2))
    dY = torch.rand(n, D_FEAT, generator=torch.Generator().manual_seed(3))
    Xn, dYn = X.numpy(), dY.numpy()

    def step_c():
        c_spmm.spmm_f32(rp, cl, Xn)
        c_spmm.spmm_f32(rpt, clt, dYn)

    def run(fn, budget):
        fn()                                   # warm-up (page faults, thread pool)
        ts = []
        t_end = time.perf_counter() + budget
        while len(ts) < max_steps and (len(ts) < 2 or time.perf_counter() < t_end):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts), len(ts)

    results, counts = {}, {}
    results["openmp_c"], counts["openmp_c"] = run(step_c, budget_s * 0.6)
    if n_edges <= 200_000_000:                 # torch's CSR tensors need int64 columns: 3x the index memory
        A = torch.sparse_csr_tensor(torch.from_numpy(rp), torch.from_numpy(cl).to(torch.int64), torch.ones(cl.size), size=(n, n))
        At = torch.sparse_csr_tensor(torch.from_numpy(rpt), torch.from_numpy(clt).to(torch.int64), torch.ones(clt.size), size=(n, n))

        def step_torch():
            A @ X
            At @ dY

        results["torch_sparse_csr"], counts["torch_sparse_csr"] = run(step_torch, budget_s * 0.4)
    best = min(results, key=results.get)
    return {
        "value": n_edges / results[best], "unit": "edges/s", "cores": cores, "kind": "port",
        "sample": f"the whole workload: all {n_edges} edges of the R-MAT stream (scale {scale}, |V|={n}), "
                  f"SpMM fwd+bwd d={D_FEAT}, {best}, median of {counts[best]} steps",
        "ms_per_step": results[best] * 1e3, "steps": counts[best],
        "variants_ms": {k: v * 1e3 for k, v in results.items()},
    }


def cpu_pubmed_leg(steps: int = 2, name: str = "pubmed"):
    """Reference train step on the Pubmed- (configs[1]) or Cora-shaped (configs[0]) graph on the host: dense adj,
    encoder, dropout, Z Z^T, weighted BCE, backward, Adam (train_transductive.py:55-68 repaired)."""
    from gae_dgl_b200 import synthetic
    from oracle import gae_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g, X = synthetic.planetoid_like(name, seed=0)
    c = g.csr()
    torch.manual_seed(0)
    model = O.OracleGAE(X.shape[1], [32, 16])
    opt = torch.optim.Adam(model.parameters(), lr=1e-2)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        adj = O.dense_adj_from_csr(c.rowptr, c.col)
        pw = O.pos_weight_transductive(adj)
        loss = O.bce_loss(model(c.rowptr, c.col, X), adj, pw)
        opt.zero_grad()
        loss.backward()
        opt.step()
        float(loss)
        ts.append(time.perf_counter() - t0)
    t = min(ts)
    return {"value": g.number_of_edges() / t, "unit": "edges/s", "cores": cores, "kind": "port",
            "sample": f"full {name.capitalize()}-shaped train step (N={g.number_of_nodes()}, E={g.number_of_edges()}, "
                      f"F={X.shape[1]}), best of {steps}", "ms_per_step": t * 1e3}


# ------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------

def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_gpus = args.gpus
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    scale, total, scaling = workload_of(args, n_gpus)
    leg = cpu_rmat_leg(scale, total, args.cpu_seconds, max_steps=max(2, min(args.steps, 10)))
    line = {
        "impl": "reference", "metric": METRIC, "value": leg["value"], "unit": "edges/s", "n_gpus": n_gpus,
        "steps": leg["steps"], "warmup": 1, "ms_per_step": leg["ms_per_step"], "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(scale, total),
        "note": "reference = CPU oracle port on the host cores (DGL not installable: no wheel, no network); "
                f"steps bounded by --cpu-seconds {args.cpu_seconds:g}, the graph is the full workload",
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": leg["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "variants_ms": leg["variants_ms"],
    }
    if not args.no_pubmed and n_gpus == 1 and not args.strong:
        line["cora"] = cpu_pubmed_leg(steps=5, name="cora")      # configs[0]: the reference's CPU-runnable case
        line["pubmed"] = cpu_pubmed_leg()
        line["zinc"] = cpu_zinc_leg()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------

def sampled_parity(rowptr, col, Y, seed, lo=0, n_local=None, halo_ids=None, n_sample=96):
    """Oracle check inside the bench run: the result rows of a sample of this rank's vertices (its heaviest
    hub rows, then evenly spaced ones) are recomputed on the host by the oracle's fp64-accumulating C loop
    from the global definition Y[v] = sum_{u->v} X[u], with X regenerated from the vertex ids
    (synthetic.hashed_normal_rows) -- independent of the exchange.  Returns max |dY| / max(|ref|, 1)."""
    from gae_dgl_b200 import synthetic
    from oracle import c_spmm
    n_rows = rowptr.numel() - 1
    n_local = n_rows if n_local is None else n_local
    deg = rowptr[1:] - rowptr[:-1]
    hubs = torch.topk(deg, min(8, n_rows)).indices
    spaced = torch.arange(0, n_rows, max(1, n_rows // (n_sample - 8)), device=rowptr.device)[: n_sample - 8]
    rows = torch.unique(torch.cat([hubs, spaced]))
    starts, ends = rowptr[rows], rowptr[rows + 1]
    lens = (ends - starts)
    sub_rowptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=rowptr.device)
    sub_rowptr[1:] = torch.cumsum(lens, 0)
    eidx = torch.repeat_interleave(starts - sub_rowptr[:-1], lens) + torch.arange(int(sub_rowptr[-1]), device=rowptr.device)
    c = col[eidx].to(torch.int64)
    if halo_ids is not None:
        gid = torch.where(c < n_local, c + lo, halo_ids[(c - n_local).clamp_(min=0)])
    else:
        gid = c
    uniq, inv = torch.unique(gid, return_inverse=True)
    Xs = synthetic.hashed_normal_rows(uniq, D_FEAT, seed).cpu().numpy()
    ref = c_spmm.spmm_f64acc(sub_rowptr.cpu().numpy(), inv.to(torch.int32).cpu().numpy(), Xs)
    got = Y[rows].double().cpu().numpy()
    err = float(np.abs(got - ref).max() / max(float(np.abs(ref).max()), 1.0))
    return err, int(rows.numel()), int(lens.max())


def cuda_time_ms(fn, steps, stream):
    """K calls bracketed by events on the launching stream; returns total ms."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1)


def decoder_roofline(dev, g, d=16, iters=20):
    """Roofline of the dominant kernel of the small-graph train steps: the fused decoder (dense pass over the pairs
    + partial-slot reduce + finalize + loss reduce), timed ALONE with CUDA events on its launching stream.
    From 5632 rows (d <= 16) the dense pass is the tcgen05 symmetric-half kernel: it evaluates N^2 / 2 pairs, so the
    algorithmic work per launch is 3 N^2 d flop (S over the upper triangle: N^2 d; G_I and G_J:
    N^2 d each) and N^2 MUFU ops (ex2 + rcp per evaluated pair; lg2 folded 32:1); below that the mma.sync / SIMT forms
    walk all N^2 pairs: 4 N^2 d flop, 2 N^2 MUFU ops.  `bound` is "tensor": peak = the MEASURED bf16 rate (fp16 operands
    run at the bf16 rate; half of it for the TF32 mma.sync form; split precision triples the issued MMAs, which the
    algorithmic count deliberately ignores).  Two tighter floors are reported beside it: the MUFU floor (16 ops/clk/SM
    x 148 SMs x max SM clock) and, for the tcgen05 form, the measured MMA-issue floor (35 M = 128 MMAs per 128 x 128
    tile at 74 / 98 / 110 cycles each whatever N is, tools/mma_bench.cu -> profiles/r02_mma_bench.log)."""
    from gae_dgl_b200 import ops
    n = g.number_of_nodes()
    c, t = g.csr(), g.csr_t()
    Zd = torch.randn(n, d, device=dev) * 0.3
    pw = 5.0
    st = torch.cuda.current_stream()

    def fn():
        ops.decoder_bce(Zd, c.rowptr, c.col, t.rowptr, t.col, pw, want_loss=True, want_grad=True)

    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = cuda_time_ms(fn, iters, st) / iters
    tc = d <= 16 and n >= 5632                      # DEC_TC_AUTO_ROWS (csrc/decoder.cu)
    flops = (3.0 if tc else 4.0) * n * n * d
    mufu = (1.0 if tc else 2.0) * n * n
    pk = {}
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            pk = json.load(f)
    bf16 = float(pk.get("bf16_tflops", 1590.0))
    sm_mhz = float(pk.get("sm_max_mhz", 1965.0))
    peak = bf16 if tc else 0.5 * bf16
    mufu_floor_ms = mufu / (16 * 148 * sm_mhz * 1e6) * 1e3
    ach = flops / (ms * 1e-3) / 1e12
    out = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
           "kernel": "gae_decoder_bce_f32 (" + ("tcgen05 fp16-split symmetric-half dense pass + slot reduce" if tc else
                                                "mma.sync TF32 dense pass" if d <= 16 else "SIMT dense pass")
                     + " + finalize + loss reduce), timed alone",
           "ms": ms, "algorithmic_flops": flops, "mufu_ops": mufu, "mufu_floor_ms": mufu_floor_ms,
           "frac_mufu": mufu_floor_ms / ms,
           "peak_source": (("" if tc else "0.5 x ") + "measured bf16 burst (MEASURED_PEAKS.json)" if pk else
                           ("" if tc else "0.5 x ") + "fallback bf16 1.59 PF")}
    if tc:
        tiles = (n + 127) // 128
        tiles = tiles * (tiles + 1) // 2
        mma_floor_ms = tiles * (3 * 110 + 16 * 74 + 16 * 98) / (148 * sm_mhz * 1e6) * 1e3
        out.update({"mma_issue_floor_ms": mma_floor_ms, "frac_mma_issue": mma_floor_ms / ms})
    return out


def pubmed_leg(dev, steps=100, warmup=5, name="pubmed"):
    """configs[1]: Pubmed-shaped transductive train step (fwd + bwd + Adam, fused decoder), as
    train_transductive.py runs it: the step captured in a CUDA graph and replayed."""
    import gae_dgl_b200 as G
    from gae_dgl_b200 import synthetic
    from gae_dgl_b200.graphed import GraphedTrainStep
    g, X = synthetic.planetoid_like(name, seed=0)
    torch.manual_seed(0)
    model = G.GAE(X.shape[1], [32, 16]).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, capturable=True, fused=True)
    g.to(dev)
    Xd = X.to(dev)
    pw = G.pos_weight_of(g, transductive=True)
    st = torch.cuda.current_stream()
    from gae_dgl_b200 import ops
    # as train_transductive.py runs it: the first layer's A X is loop-invariant and computed once per change of
    # the features (bit-identical); in the e2e variant the features arrive from the host EVERY step, so there
    # the aggregation is part of the step again
    agg = ops.spmm(g.csr().rowptr, g.csr().col, Xd, g.csr().plan)

    def loss_fn():
        g.ndata["h"] = agg
        return model.loss(g, pos_weight=pw, aggregated_input=True)

    def loss_fn_e2e():
        g.ndata["h"] = Xd
        return model.loss(g, pos_weight=pw)

    step = GraphedTrainStep(model, opt, loss_fn, warmup=warmup)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ms = cuda_time_ms(step, steps, st) / steps
    # e2e: features from pinned host memory every step, loss read back every step
    Xp = X.pin_memory()
    step2 = GraphedTrainStep(model, opt, loss_fn_e2e, warmup=warmup)

    def step_e2e():
        Xd.copy_(Xp, non_blocking=True)
        return step2().item()

    step_e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        last = step_e2e()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    e = g.number_of_edges()
    return {"workload": f"{name}_like_N{g.number_of_nodes()}_E{e}_F{X.shape[1]} train step (fwd+bwd+Adam, fused decoder, "
                        "CUDA-graph replay, input aggregation hoisted out of the epoch loop)",
            "value": e / (ms * 1e-3), "unit": "edges/s", "ms_per_step": ms, "final_loss": last,
            "roofline": decoder_roofline(dev, g),
            "e2e": {"value": e / (ms_e2e * 1e-3), "unit": "edges/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(X.numel() * 4), "d2h_bytes_per_step": 4}}


def zinc_leg(dev, steps=60, warmup=5, batch_size=256):
    """configs[2]: ZINC-shaped inductive step, batch = 256 molecules, hidden 32/16, as train_inductive.py runs it:
    device collation from the packed dataset + the native train step (gae_step_fwd_bwd_f32 + gae_adam_step_f32),
    a new random batch every step.  `value`: losses stay on the device (one sync at the end); `e2e`: the batch's
    molecule ids come from pinned host memory and the loss is read back EVERY step (train_inductive.py:53)."""
    import gae_dgl_b200 as G
    from gae_dgl_b200 import synthetic
    from gae_dgl_b200.graph import PackedGraphDataset
    from gae_dgl_b200.native_step import NativeTrainStep
    ds = synthetic.zinc_like_dataset(4096, seed=0)
    packed = PackedGraphDataset(ds, dev)
    torch.manual_seed(0)
    model = G.GAE(39, [32, 16]).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    native = NativeTrainStep(model, opt)
    rng = np.random.default_rng(0)
    batches = [rng.permutation(len(ds))[:batch_size] for _ in range(2 * steps + warmup)]
    edges = float(np.mean([packed.edges[b].sum() for b in batches[warmup:]]))

    def step(ids):
        return native(packed.batch(ids))

    for ids in batches[:warmup]:
        step(ids)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for ids in batches[warmup:warmup + steps]:
        last = step(ids)
    last = float(last)                       # one sync at the end; includes host collation cost
    ms = (time.perf_counter() - t0) * 1e3 / steps
    t0 = time.perf_counter()
    for ids in batches[warmup + steps:]:
        last_e2e = float(step(ids))          # device -> host read of the loss every step
    ms_e2e = (time.perf_counter() - t0) * 1e3 / steps
    big = G.batch(ds[:batch_size], device=dev)
    return {"workload": f"zinc_like batch={batch_size} (mean {edges:.0f} directed edges/batch) inductive train step: "
                        "device collation + native fwd/bwd + native Adam, wall clock incl. host overhead",
            "value": edges / (ms * 1e-3), "unit": "edges/s", "ms_per_step": ms, "final_loss": last,
            "e2e": {"value": edges / (ms_e2e * 1e-3), "unit": "edges/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(batch_size * 8 + 2 * (batch_size + 1) * 8), "d2h_bytes_per_step": 4,
                    "final_loss": last_e2e,
                    "api": "PackedGraphDataset.batch(host ids) + NativeTrainStep; loss.item() every step"},
            "roofline": decoder_roofline(dev, big)}


def cpu_zinc_leg(steps=2, batch_size=256):
    """Reference inductive step on the host (train_inductive.py:43-53): dgl.batch, dense adj, forward, BCE,
    backward, Adam."""
    from gae_dgl_b200 import synthetic
    from oracle import gae_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ds = synthetic.zinc_like_dataset(batch_size * (steps + 1), seed=0)
    torch.manual_seed(0)
    model = O.OracleGAE(39, [32, 16])
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    ts, edges = [], []
    for k in range(steps + 1):
        part = ds[k * batch_size:(k + 1) * batch_size]
        t0 = time.perf_counter()
        s, d, n = O.batch_graphs([(*g.edges(), g.number_of_nodes()) for g in part])
        rowptr, col = O.coo_to_csr(s, d, n)
        X = torch.cat([g.ndata["h"] for g in part])
        adj = O.dense_adj(s, d, n)
        pw = O.pos_weight_inductive(adj)
        loss = O.bce_loss(model(rowptr, col, X), adj, pw)
        opt.zero_grad()
        loss.backward()
        opt.step()
        float(loss)
        if k:
            ts.append(time.perf_counter() - t0)
            edges.append(s.numel())
    t = min(ts)
    return {"value": float(np.mean(edges)) / t, "unit": "edges/s", "cores": cores, "kind": "port",
            "sample": f"batch={batch_size} ZINC-shaped train step incl. collation, best of {steps}", "ms_per_step": t * 1e3}


def run_ours(args):
    import torch.distributed as dist
    import gae_dgl_b200 as G
    from gae_dgl_b200 import _lib, ops, synthetic
    from gae_dgl_b200.graph import coo_to_csr_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    for kv in filter(None, args.tune.split(",")):
        k, v = kv.split("=")
        _lib.set_tuning(k, int(v))

    scale, total_edges, scaling = workload_of(args, world)
    n = 1 << scale
    stream = torch.cuda.current_stream()

    if world == 1:
        src, dst = synthetic.rmat_edges(scale, total_edges, seed=1, device=dev)
        rowptr, col = coo_to_csr_torch(src, dst, n)
        rowptr_t, col_t = coo_to_csr_torch(dst, src, n)
        del src, dst
        torch.cuda.empty_cache()
        g = G.DGLGraph.from_csr(rowptr, col, csr_t=(rowptr_t, col_t))
        c, t = g.csr(), g.csr_t()
        X = synthetic.hashed_normal(n, D_FEAT, 2, device=dev)
        dY = synthetic.hashed_normal(n, D_FEAT, 3, device=dev)
        Y = torch.empty_like(X)
        dX = torch.empty_like(X)
        ws_f = c.plan.workspace(D_FEAT, dev)
        ws_b = t.plan.workspace(D_FEAT, dev)
        local_edges = total_edges
        local_rows = n

        def fwd():
            ops.spmm(c.rowptr, c.col, X, c.plan, out=Y, partial_ws=ws_f)

        def bwd():
            ops.spmm(t.rowptr, t.col, dY, t.plan, out=dX, partial_ws=ws_b)

        exchange_desc = "none (single GPU)"
        halo_rows = 0
        halo_stats = None
    else:
        from gae_dgl_b200 import parallel
        part = parallel.build_rmat_partition(scale, total_edges, seed=1, d=D_FEAT, device=dev,
                                             exchange=args.exchange, stages=args.stages, push_ctas=args.push_ctas,
                                             kind=args.halo_kind,
                                             thresholds=tuple(int(x) for x in args.hot.split(",")) if args.hot else None)
        fwd, bwd = part.fwd, part.bwd
        local_edges, local_rows = part.local_edges, part.local_rows
        exchange_desc = part.exchange_desc
        halo_rows = part.halo_rows
        halo_stats = getattr(getattr(part.fwd_op, "sp", None), "stats", None)

    def step():
        fwd()
        bwd()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, barrier + synchronize on both sides, CUDA events, max over ranks
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    ev[0].record(stream)
    for i in range(args.steps):
        fwd()
        ev[2 * i + 1].record(stream)
        bwd()
        ev[2 * i + 2].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop()
    launches = _lib.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    fwd_ms = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(args.steps)]
    bwd_ms = [ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(args.steps)]
    tmax = torch.tensor([total_ms, statistics.mean(fwd_ms), statistics.mean(bwd_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms, fwd_mean, bwd_mean = (float(x) for x in tmax)
    ms_per_step = total_ms / args.steps
    value = total_edges / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (SpMM forward), per rank
    peak, peak_src = peaks()
    alg_bytes = spmm_algorithmic_bytes(local_edges, local_rows, D_FEAT)
    # the fwd event pair brackets the row kernel + hub-segment kernel + hub reduce (+ exchange at N>1)
    achieved = alg_bytes / (statistics.mean(fwd_ms) * 1e-3) / 1e9
    traffic = traffic_bwd = None
    tpath = os.path.join(ROOT, "profiles", "spmm_traffic.json")
    if os.path.exists(tpath) and world == 1 and (scale, total_edges) == (22, 100_000_000):
        # an ncu capture of the single-GPU C4 launch (profiles/): means nothing for any other workload
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_bwd = tj.get("dram_bytes_per_launch"), tj.get("bwd_dram_bytes_per_launch")
    # SURVEY 8(d) secondary model: every byte touched once (col ids, row pointers, X read once, Y written once)
    compulsory = 4 * local_edges + 8 * (local_rows + 1) + 2 * 4 * D_FEAT * local_rows
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "spmm_fused_kernel (fwd: every row class + hub segments in one launch) + spmm_hub_reduce_kernel",
                "algorithmic_bytes": alg_bytes, "fwd_ms": statistics.mean(fwd_ms), "bwd_ms": statistics.mean(bwd_ms),
                "peak_source": peak_src, "compulsory_bytes": compulsory,
                "frac_compulsory": compulsory / (statistics.mean(fwd_ms) * 1e-3) / 1e9 / peak}
    if traffic_bwd:
        roofline["traffic_bwd"] = traffic_bwd
    if traffic and world == 1:
        # the DRAM-side view of the same launch: bytes that actually crossed HBM (ncu capture of this
        # workload, profiles/spmm_traffic.json) over the live forward time; `frac` above can exceed 1
        # because gathered source rows are re-read from L2, this one cannot
        roofline["achieved_dram"] = traffic / (statistics.mean(fwd_ms) * 1e-3) / 1e9
        roofline["frac_dram"] = roofline["achieved_dram"] / peak

    line = {
        "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(scale, total_edges),
        "impl_detail": {"partition": "1d_vertex_blocks" if world > 1 else "single", "exchange": exchange_desc,
                        "halo_rows_rank0": halo_rows, "halo_exchange_rank0": halo_stats, "local_rows_rank0": local_rows, "local_edges_rank0": local_edges,
                        "tuning": {k: _lib.get_tuning(k) for k in ("spmm_unroll", "spmm_block", "spmm_cache",
                                                                   "spmm_rows_per_warp", "spmm_fused")}},
        "roofline": roofline, "gpu_launches": int(launches), "clocks": sampler.summary(),
        "fwd_ms_max_over_ranks": fwd_mean, "bwd_ms_max_over_ranks": bwd_mean,
    }

    # ---- parity, in the same run: sampled result rows of every rank against the host oracle
    if world == 1:
        ef = sampled_parity(c.rowptr, c.col, Y, 2)
        eb = sampled_parity(t.rowptr, t.col, dX, 3)
    else:
        for op in (part.fwd_op, part.bwd_op):
            if hasattr(op, "check"):
                op.check()                       # raises if any device-side flag wait timed out
        lo = part.fwd_op.hp.bounds[rank]
        ef = sampled_parity(part.fwd_op.hp.rowptr, part.fwd_op.hp.col, part.fwd_op.Y, 2, lo, part.fwd_op.hp.n_local,
                            part.fwd_op.hp.halo_ids)
        eb = sampled_parity(part.bwd_op.hp.rowptr, part.bwd_op.hp.col, part.bwd_op.Y, 3, lo, part.bwd_op.hp.n_local,
                            part.bwd_op.hp.halo_ids)
    perr = torch.tensor([ef[0], eb[0]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(perr, op=dist.ReduceOp.MAX)
    line["parity"] = {"fwd_max_err": float(perr[0]), "bwd_max_err": float(perr[1]), "tolerance": 1e-5,
                      "ok": bool(perr.max() < 1e-5), "rows_per_rank": ef[1], "max_degree_checked_rank0": ef[2],
                      "checker": "oracle/spmm_ref.c fp64-accumulating loop on sampled rows of every rank (hub rows included), "
                                 "X regenerated from vertex ids; max over ranks"}

    # ---- e2e: host buffers through the C ABI entry point, copies inside the timed region
    if not args.no_e2e and world == 1:
        line["e2e"] = e2e_leg(c, t, X, dY, n, total_edges, min(args.steps, 20), dev, ws_f, ws_b)
    elif not args.no_e2e:
        line["e2e"] = part.e2e(min(args.steps, 10))

    if rank == 0 and world == 1 and not args.no_cpu:
        del X, dY
        torch.cuda.empty_cache()
        del Y, dX, c, t, g, fwd, bwd, step
        torch.cuda.empty_cache()
        leg = cpu_rmat_leg(scale, total_edges, min(args.cpu_seconds, 20.0), max_steps=5)
        line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if not args.no_pubmed and not args.strong:
            line["cora"] = pubmed_leg(dev, name="cora")           # configs[0] (CPU-runnable reference case) beside its CPU time
            line["cora"]["cpu_baseline"] = cpu_pubmed_leg(steps=5, name="cora")
            line["pubmed"] = pubmed_leg(dev)
            line["pubmed"]["cpu_baseline"] = cpu_pubmed_leg()
            line["zinc"] = zinc_leg(dev)
            line["zinc"]["cpu_baseline"] = cpu_zinc_leg()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def e2e_leg(c, t, X, dY, n, total_edges, steps, dev, ws_f, ws_b):
    """Same step through gae_spmm_csr_f32_host: features / gradients live in PINNED HOST memory,
    results are returned to pinned host memory; the graph (CSR, CSR^T, hub plans) is resident
    device state, as the DGLGraph is across epochs in the reference.

    Every call is H2D -> kernels -> D2H in stream order.  The forward and the backward call of a
    step are independent requests and run on two streams; consecutive steps alternate between two
    sets of (streams, staging buffers, host result buffers), so that the D2H of one step overlaps the
    H2D of the next (PCIe is full duplex) instead of queueing behind it in the same stream."""
    import ctypes
    from gae_dgl_b200 import _lib
    lib = _lib.load()
    pinned = lambda: torch.empty((n, D_FEAT), dtype=torch.float32, pin_memory=True)  # noqa: E731
    Xh, dYh = pinned(), pinned()
    Xh.copy_(X)
    dYh.copy_(dY)
    cur = torch.cuda.current_stream()
    p = lambda x: ctypes.c_void_p(x.data_ptr()) if x is not None else None  # noqa: E731

    class Lane:
        """One in-flight request slot: stream, device staging, host result buffer, hub workspace."""
        def __init__(self, csr, src_h, ws):
            self.csr, self.src_h = csr, src_h
            self.st = torch.cuda.Stream()
            self.xs, self.ys = torch.empty_like(X), torch.empty_like(X)
            self.out_h = pinned()
            self.ws = ws if ws is None else torch.empty_like(ws)

        def call(self):
            rc = lib.gae_spmm_csr_f32_host(p(self.csr.rowptr), p(self.csr.col), p(self.src_h), n, D_FEAT, p(self.out_h),
                                           D_FEAT, n, D_FEAT, ctypes.byref(self.csr.plan.struct), p(self.ws), p(self.xs),
                                           p(self.ys), self.st.cuda_stream)
            _lib.check(rc, "gae_spmm_csr_f32_host")

    sets = [(Lane(c, Xh, ws_f), Lane(t, dYh, ws_b)) for _ in range(2)]
    lanes = [ln for pair in sets for ln in pair]

    def timed(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for ln in lanes:
            ln.st.wait_event(e0)
        for i in range(k):
            for ln in sets[i % 2]:
                ln.call()
        for ln in lanes:
            cur.wait_stream(ln.st)
        e1.record(cur)
        e1.synchronize()
        return e0.elapsed_time(e1)

    timed(2)
    torch.cuda.synchronize()
    ms = timed(steps) / steps
    ok = all(bool(torch.isfinite(ln.out_h[:1024]).all()) for ln in lanes)
    same = bool(torch.equal(sets[0][0].out_h[:4096], sets[1][0].out_h[:4096]))
    return {"value": total_edges / (ms * 1e-3), "unit": "edges/s", "ms_per_step": ms, "steps": steps,
            "h2d_bytes_per_step": int(2 * n * D_FEAT * 4), "d2h_bytes_per_step": int(2 * n * D_FEAT * 4),
            "api": "gae_spmm_csr_f32_host (C ABI), pinned host X/dY in, Y/dX out; graph resident; forward and backward "
                   "call on two streams, consecutive steps on alternating stream/buffer sets (D2H of step k overlaps H2D of k+1)",
            "finite": ok, "sets_agree": same}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
