"""Host-side multi-GPU logic on CPU: ciphertext sharding, key broadcast and result gathering, run as
world_size-2 gloo jobs (no GPU, no compute entry point of the CUDA library is called)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mosfhet_b200 import sharding  # noqa: E402


def test_shard_bounds_cover_and_balance():
    for count in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            b = sharding.shard_bounds(count, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == count
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
            assert sharding.my_shard(count, world - 1, world) == b[-1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmpdir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. key broadcast: rank 0 owns the "resident" buffers, everyone ends with identical copies
        n = 1 << 18
        bsk = torch.arange(n, dtype=torch.float64) * 0.5 if rank == 0 else torch.zeros(n, dtype=torch.float64)
        ksk = torch.arange(n, dtype=torch.int64) * 3 + 1 if rank == 0 else torch.zeros(n, dtype=torch.int64)
        sharding.broadcast_buffer(bsk, 0, chunk_bytes=1 << 16)        # many chunks
        sharding.broadcast_buffer(ksk, 0, chunk_bytes=1 << 20)
        assert torch.equal(bsk, torch.arange(n, dtype=torch.float64) * 0.5)
        assert torch.equal(ksk, torch.arange(n, dtype=torch.int64) * 3 + 1)
        # 2. shard -> "process" -> gather: rank 0 sees every ciphertext's result in input order
        count, width = 1001, 5                                         # ragged: 501 + 500
        b, e = sharding.my_shard(count, rank, world)
        full = (np.arange(count * width, dtype=np.uint64).reshape(count, width) * np.uint64(0x9E3779B97F4A7C15))
        local = full[b:e] ^ np.uint64(0xFFFF)                          # stand-in for the per-rank PBS+KS
        got = sharding.gather_results(local, count)
        if rank == 0:
            assert got.shape == (count, width)
            assert np.array_equal(got, full ^ np.uint64(0xFFFF))
        else:
            assert got is None
        open(os.path.join(tmpdir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_broadcast_and_gather_world2(tmp_path):
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
