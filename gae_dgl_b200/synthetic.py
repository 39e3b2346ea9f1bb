"""Shape-faithful synthetic stand-ins for the reference's datasets (none are on disk and there
is no network; SURVEY.md section 8d gives the shapes) and the Graph500 R-MAT generator.

Every generator is a pure function of its seed and produces IDENTICAL output on CPU and on
CUDA (counter-based integer hashing instead of device RNG streams), so the CPU baseline and the
GPU path see the same graph, and a "sample" of a workload is an exact prefix of its edge list.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from .graph import DGLGraph

_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    """Python int -> two's-complement int64 value."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z: torch.Tensor, k: int) -> torch.Tensor:
    return (z >> k) & ((1 << (64 - k)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    """SplitMix64 finaliser on int64 tensors (wrapping arithmetic; same bits on CPU and CUDA)."""
    z = x + _s64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def scramble(v: torch.Tensor, scale: int, seed: int) -> torch.Tensor:
    """Bijective relabelling of [0, 2^scale) (Graph500 permutes vertex labels so that degree
    does not correlate with id): rounds of odd-multiply and xor-shift modulo 2^scale."""
    mask = (1 << scale) - 1
    k1 = (splitmix_scalar(seed * 2 + 1) | 1) & mask
    k2 = (splitmix_scalar(seed * 2 + 2) | 1) & mask
    h = max(scale // 2, 1)
    v = (v * k1) & mask
    v = v ^ (v >> h)
    v = (v * k2) & mask
    v = v ^ (v >> h)
    return v


def splitmix_scalar(x: int) -> int:
    z = (x + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def rmat_edges(scale: int, n_edges: int, seed: int = 1, device="cpu", a=0.57, b=0.19, c=0.19,
               permute: bool = True, first_edge: int = 0, chunk: int = 1 << 24) -> Tuple[torch.Tensor, torch.Tensor]:
    """Graph500 R-MAT: edges first_edge .. first_edge+n_edges of the stream defined by `seed`
    (no dedup, self loops kept -- DGL multigraph semantics).  Returns (src, dst) int64.
    Each 64-bit hash supplies four 16-bit quadrant draws."""
    device = torch.device(device)
    ta = int(round(a * 65536))
    tab = int(round((a + b) * 65536))
    tabc = int(round((a + b + c) * 65536))
    groups = (scale + 3) // 4
    srcs, dsts = [], []
    for lo in range(first_edge, first_edge + n_edges, chunk):
        m = min(chunk, first_edge + n_edges - lo)
        e = torch.arange(lo, lo + m, dtype=torch.int64, device=device)
        src = torch.zeros(m, dtype=torch.int64, device=device)
        dst = torch.zeros(m, dtype=torch.int64, device=device)
        level = 0
        for gidx in range(groups):
            h = splitmix64(e * groups + gidx + _s64(splitmix_scalar(seed)))
            for q in range(4):
                if level >= scale:
                    break
                r = _lsr(h, 16 * q) & 0xFFFF
                sbit = (r >= tab).to(torch.int64)                       # c or d quadrant
                dbit = (((r >= ta) & (r < tab)) | (r >= tabc)).to(torch.int64)   # b or d quadrant
                src = (src << 1) | sbit
                dst = (dst << 1) | dbit
                level += 1
        if permute:
            src = scramble(src, scale, seed)
            dst = scramble(dst, scale, seed)
        srcs.append(src)
        dsts.append(dst)
    if len(srcs) == 1:
        return srcs[0], dsts[0]
    return torch.cat(srcs), torch.cat(dsts)


def hashed_normal(n_rows: int, d: int, seed: int, device="cpu", first_row: int = 0) -> torch.Tensor:
    """Deterministic ~N(0,1) fp32 features [n_rows, d]: sum of four hashed uniforms (Irwin-Hall),
    identical on every device; row r depends only on (seed, first_row + r)."""
    device = torch.device(device)
    out = torch.empty((n_rows, d), dtype=torch.float32, device=device)
    rows_per = max(1, (1 << 24) // max(d, 1))
    for lo in range(0, n_rows, rows_per):
        m = min(rows_per, n_rows - lo)
        idx = (torch.arange(lo + first_row, lo + first_row + m, dtype=torch.int64, device=device)[:, None] * d
               + torch.arange(d, dtype=torch.int64, device=device)[None, :])
        h = splitmix64(idx + _s64(splitmix_scalar(seed + 77)))
        s = torch.zeros((m, d), dtype=torch.float32, device=device)
        for q in range(4):
            s += (_lsr(h, 16 * q) & 0xFFFF).to(torch.float32)
        # sum of 4 U{0..65535}: mean 2*65535, var 4*(65536^2-1)/12
        out[lo:lo + m] = (s - 2.0 * 65535.0) / float(np.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0))
    return out


def hashed_normal_rows(ids: torch.Tensor, d: int, seed: int) -> torch.Tensor:
    """Rows `ids` (int64, any order) of the matrix hashed_normal(n, d, seed) without building it."""
    idx = ids.to(torch.int64)[:, None] * d + torch.arange(d, dtype=torch.int64, device=ids.device)[None, :]
    h = splitmix64(idx + _s64(splitmix_scalar(seed + 77)))
    s = torch.zeros(idx.shape, dtype=torch.float32, device=ids.device)
    for q in range(4):
        s += (_lsr(h, 16 * q) & 0xFFFF).to(torch.float32)
    return (s - 2.0 * 65535.0) / float(np.sqrt(4.0 * (65536.0 ** 2 - 1.0) / 12.0))


# ---- Planetoid-shaped citation graphs ------------------------------------------------------------

PLANETOID_SHAPES = {
    # name: (N, undirected edges, self loops, F, nnz per feature row)
    "cora": (2708, 5278, 0, 1433, 18),
    "citeseer": (3327, 4552, 0, 3703, 32),
    "pubmed": (19717, 44325, 1, 500, 50),
}


def planetoid_like(name: str = "cora", seed: int = 0):
    """(DGLGraph, features fp32 [N,F]) with the node / directed-edge / feature counts of the
    Planetoid dataset `name`: random undirected edges mirrored to both directions
    (citation graphs are symmetric), row-normalised sparse non-negative features."""
    n, und, loops, f, nnz = PLANETOID_SHAPES[name.lower()]
    rng = np.random.default_rng(seed)
    pairs = set()
    # preferential-attachment-flavoured endpoints so the degree distribution has a tail
    weights = 1.0 / np.sqrt(np.arange(1, n + 1))
    weights /= weights.sum()
    while len(pairs) < und:
        need = und - len(pairs)
        u = rng.choice(n, size=need * 2, p=weights)
        v = rng.integers(0, n, size=need * 2)
        for x, y in zip(u.tolist(), v.tolist()):
            if x != y:
                pairs.add((min(x, y), max(x, y)))
                if len(pairs) >= und:
                    break
    pairs = np.asarray(sorted(pairs), dtype=np.int64)
    perm = rng.permutation(n)
    s, d = perm[pairs[:, 0]], perm[pairs[:, 1]]
    src = np.concatenate([s, d, np.arange(loops)])
    dst = np.concatenate([d, s, np.arange(loops)])
    g = DGLGraph()
    g.add_nodes(n)
    g.add_edges(src, dst)
    feats = np.zeros((n, f), dtype=np.float32)
    cols = rng.integers(0, f, size=(n, nnz))
    vals = rng.random((n, nnz)).astype(np.float32) + 0.1
    np.put_along_axis(feats, cols, vals, axis=1)
    feats /= feats.sum(1, keepdims=True)
    return g, torch.from_numpy(feats)


# ---- ZINC-shaped molecular graphs ------------------------------------------------------------------

ATOM_FDIM = 39  # prepare_data.py:14-16: 23 elements + 6 degrees + 5 charges + 4 chiralities + aromatic


def zinc_like_molecule(rng: np.random.Generator) -> DGLGraph:
    """One molecule-shaped graph (prepare_data.py:38-69): n ~ clip(N(23.2, 4.5), 6, 38) heavy
    atoms, a random spanning tree plus ~Poisson(2.7) ring-closing bonds, max degree 4, every bond
    added in both directions (prepare_data.py:61-64), 39-d one-hot atom features in ndata['h']."""
    n = int(np.clip(round(rng.normal(23.2, 4.5)), 6, 38))
    deg = np.zeros(n, dtype=np.int64)
    bonds = set()
    for v in range(1, n):
        cand = np.flatnonzero(deg[:v] < 3)
        u = int(rng.choice(cand)) if cand.size else int(np.argmin(deg[:v]))
        bonds.add((u, v))
        deg[u] += 1
        deg[v] += 1
    for _ in range(int(rng.poisson(2.7))):
        cand = np.flatnonzero(deg < 4)
        if cand.size < 2:
            break
        u, v = rng.choice(cand, size=2, replace=False)
        u, v = int(min(u, v)), int(max(u, v))
        if (u, v) in bonds:
            continue
        bonds.add((u, v))
        deg[u] += 1
        deg[v] += 1
    b = np.asarray(sorted(bonds), dtype=np.int64)
    feats = np.zeros((n, ATOM_FDIM), dtype=np.float32)
    elem = rng.choice(23, size=n, p=_ELEM_P)
    feats[np.arange(n), elem] = 1.0
    feats[np.arange(n), 23 + np.minimum(deg, 5)] = 1.0
    feats[np.arange(n), 29 + rng.choice(5, size=n, p=[0.02, 0.03, 0.9, 0.03, 0.02])] = 1.0
    feats[np.arange(n), 34 + rng.choice(4, size=n, p=[0.85, 0.07, 0.07, 0.01])] = 1.0
    feats[:, 38] = rng.random(n) < 0.4
    g = DGLGraph()
    g.add_nodes(n)
    g.add_edges(np.concatenate([b[:, 0], b[:, 1]]), np.concatenate([b[:, 1], b[:, 0]]))
    g.ndata['h'] = torch.from_numpy(feats)
    return g


_ELEM_P = np.asarray([0.72, 0.10, 0.11] + [0.07 / 20] * 20)
_ELEM_P = _ELEM_P / _ELEM_P.sum()


def zinc_like_dataset(n_graphs: int, seed: int = 0):
    rng = np.random.default_rng(seed)
    return [zinc_like_molecule(rng) for _ in range(n_graphs)]
