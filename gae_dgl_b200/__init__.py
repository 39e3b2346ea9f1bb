"""gae_dgl_b200 -- B200-native GAE encoder/decoder hot path behind the gae-dgl module surface.

    from gae_dgl_b200 import GAE, VGAE, DGLGraph, batch

Importing the package does not load the CUDA library; the first op does, and raises if
`lib/libgae_b200.so` has not been built (`python -m gae_dgl_b200.build`).
"""
from .graph import DGLGraph, batch, function, init  # noqa: F401
from .gae import GAE, VGAE, GCN, NodeApplyModule, InnerProductDecoder, pos_weight_of  # noqa: F401
from .dataset import MolDataset  # noqa: F401
from ._lib import GaeError  # noqa: F401

__version__ = "0.1.0"
