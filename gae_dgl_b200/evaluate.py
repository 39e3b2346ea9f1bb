"""Link-prediction evaluation (ROC-AUC / average precision), the metric of the GAE paper the
reference cites (README.md:58) but does not implement.

Scores are sigmoid(<z_i, z_j>) on held-out positive edges and an equal number of sampled
non-edges.  Evaluation utility only: not on the training hot path."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch


def sample_non_edges(g, count: int, seed: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """`count` node pairs (i != j) that are not edges of g in either direction."""
    n = g.number_of_nodes()
    src, dst = g.edges()
    present = set((src * n + dst).tolist()) | set((dst * n + src).tolist())
    rng = np.random.default_rng(seed)
    out_i, out_j = [], []
    while len(out_i) < count:
        i = rng.integers(0, n, size=2 * (count - len(out_i)) + 16)
        j = rng.integers(0, n, size=i.size)
        for a, b in zip(i.tolist(), j.tolist()):
            if a != b and (a * n + b) not in present:
                out_i.append(a)
                out_j.append(b)
                if len(out_i) >= count:
                    break
    return np.asarray(out_i, dtype=np.int64), np.asarray(out_j, dtype=np.int64)


@torch.no_grad()
def edge_scores(z: torch.Tensor, i, j) -> torch.Tensor:
    i = torch.as_tensor(i, device=z.device, dtype=torch.int64)
    j = torch.as_tensor(j, device=z.device, dtype=torch.int64)
    return torch.sigmoid((z[i] * z[j]).sum(1))


@torch.no_grad()
def link_prediction_metrics(z: torch.Tensor, pos_edges, neg_edges):
    """-> (roc_auc, average_precision) for embeddings z [N, d]."""
    from sklearn.metrics import average_precision_score, roc_auc_score
    ps = edge_scores(z, *pos_edges).cpu().numpy()
    ns = edge_scores(z, *neg_edges).cpu().numpy()
    y = np.concatenate([np.ones_like(ps), np.zeros_like(ns)])
    s = np.concatenate([ps, ns])
    return float(roc_auc_score(y, s)), float(average_precision_score(y, s))
