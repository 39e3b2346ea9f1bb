"""CPU: the N>1 path (routing, halo plan, all-to-all-v exchange) with world_size 2 on gloo."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from gae_dgl_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_block_bounds():
    assert parallel.block_bounds(10, 4) == [0, 3, 6, 9, 10]
    assert parallel.block_bounds(8, 2) == [0, 4, 8]


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_spmm_matches_global(tmp_path, world):
    from tests import _dist_worker
    mp.spawn(_dist_worker.run, args=(world, _free_port(), 10, 20000, 8, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert os.path.exists(tmp_path / f"ok_{r}")
