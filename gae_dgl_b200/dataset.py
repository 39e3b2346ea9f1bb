"""Map-style dataset over a list of graphs -- the role of gae_dgl/dataset.py:3-12 (`MolDataset`).

The reference hands its `MolDataset` to a `DataLoader` whose `collate_fn` batches the member
graphs on the host (train_inductive.py:31-35,84).  The same class name and constructor are kept so
that script reads unchanged; what is added is what the device-side collation needs: the per-graph
node / edge counts (batches are sized without touching the members) and sequence-style access
(slices, index arrays) for building `graph.PackedGraphDataset` splits.
"""
from __future__ import annotations

from typing import Iterator, Sequence

import numpy as np
import torch.utils.data


class MolDataset(torch.utils.data.Dataset):
    """`MolDataset(graphs)`: `len(ds)`, `ds[i]`, `ds.graphs` as in the reference; plus `ds[a:b]`,
    `ds[[i, j, ...]]`, iteration, and lazily computed `num_nodes` / `num_edges` vectors."""

    def __init__(self, graphs: Sequence):
        self.graphs = list(graphs)
        self._counts = None
        print("MolDataset: {:d} graphs".format(len(self.graphs)))

    def __len__(self) -> int:
        return len(self.graphs)

    def __getitem__(self, item):
        if isinstance(item, slice):
            return self.graphs[item]
        if isinstance(item, (list, tuple, np.ndarray)):
            return [self.graphs[int(i)] for i in item]
        return self.graphs[int(item)]

    def __iter__(self) -> Iterator:
        return iter(self.graphs)

    def _count(self):
        if self._counts is None:
            self._counts = (np.fromiter((g.number_of_nodes() for g in self.graphs), dtype=np.int64, count=len(self.graphs)),
                            np.fromiter((g.number_of_edges() for g in self.graphs), dtype=np.int64, count=len(self.graphs)))
        return self._counts

    @property
    def num_nodes(self) -> np.ndarray:
        """int64 [len(ds)] nodes per graph."""
        return self._count()[0]

    @property
    def num_edges(self) -> np.ndarray:
        """int64 [len(ds)] directed edges per graph."""
        return self._count()[1]
