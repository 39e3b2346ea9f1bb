"""Loader for tests/golden/ref_gae_steps.npz -- outputs of the REFERENCE'S OWN gae.py and
Trainer.iteration (train_inductive.py:37-57) executed over a minimal DGL stand-in by
tests/golden/make_golden_reference.py.  Shared by the oracle pins (CPU) and the GPU parity tests."""
import os
from types import SimpleNamespace

import numpy as np
import torch

PATH = os.path.join(os.path.dirname(__file__), "golden", "ref_gae_steps.npz")
CASES = ("A", "B", "C", "D")
_cache = {}


def _npz():
    if "z" not in _cache:
        _cache["z"] = np.load(PATH)
    return _cache["z"]


def load_case(tag: str) -> SimpleNamespace:
    z = _npz()
    t = lambda k: torch.from_numpy(z[f"{tag}_{k}"])  # noqa: E731
    sizes = [int(s) for s in z[f"{tag}_sizes"]]
    members = [(t(f"src{i}"), t(f"dst{i}"), n, t(f"X{i}")) for i, n in enumerate(sizes)]
    hidden = [int(h) for h in z[f"{tag}_hidden"]]
    n_steps = len(z[f"{tag}_losses"])
    keys = [k[len(f"{tag}_init."):] for k in z.files if k.startswith(f"{tag}_init.")]
    state = lambda prefix: {k: t(f"{prefix}.{k}") for k in keys}  # noqa: E731
    return SimpleNamespace(
        tag=tag, members=members, sizes=sizes, n=sum(sizes), in_dim=members[0][3].shape[1], hidden=hidden,
        lr=float(z[f"{tag}_lr"]), X=torch.cat([m[3] for m in members]),
        init=state("init"), after=[state(f"after{s}") for s in range(n_steps)],
        grads=[state(f"grad{s}") for s in range(n_steps)], masks=[t(f"mask{s}") for s in range(n_steps)],
        losses=[float(x) for x in z[f"{tag}_losses"]], mask_eval=t("mask_eval"), loss_eval=float(z[f"{tag}_loss_eval"]),
        logits=t("logits"), emb=t("emb"), encode=t("encode"), adj=t("adj"), pos_weight=float(z[f"{tag}_pos_weight"]),
        in_deg=t("in_deg"))


def weights_of(state: dict, n_layers: int):
    return [(state[f"layers.{i}.apply_mod.linear.weight"], state[f"layers.{i}.apply_mod.linear.bias"])
            for i in range(n_layers)]
