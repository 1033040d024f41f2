"""Multi-GPU plumbing: independent ciphertexts are sharded across ranks, the bootstrapping and
key-switching keys are broadcast once from rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests);
there is no collective in the steady state.

One process per GPU (``torch.distributed``); everything here is host logic plus two broadcasts.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_bounds(count: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [begin, end) ranges: the first ``count % world`` ranks get one extra."""
    base, extra = divmod(count, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def my_shard(count: int, rank: int, world: int) -> Tuple[int, int]:
    return shard_bounds(count, world)[rank]


def broadcast_buffer(t, src: int = 0, chunk_bytes: int = 1 << 28):
    """In-place broadcast of a (possibly > 2 GiB) tensor from ``src`` in bounded chunks."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    flat = t.reshape(-1)
    per = max(1, chunk_bytes // flat.element_size())
    for off in range(0, flat.numel(), per):
        dist.broadcast(flat[off: off + per], src=src)
    return t


def broadcast_keys(params, bsk, ksk, device):
    """Rank 0 holds resident keys (``api.BootstrapKey`` / ``api.KeySwitchKey``); every rank returns
    key objects backed by its own HBM copy.  Non-zero ranks pass ``None`` for both keys."""
    import torch
    import torch.distributed as dist

    from . import api

    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return bsk, ksk
    rank = dist.get_rank()
    p = params.c()
    import ctypes as C
    bsk_bytes = int(api.lib().mb200_bsk_device_bytes(C.byref(p)))
    ksk_bytes = int(api.lib().mb200_ksk_device_bytes(C.byref(p)))
    bsk_buf = torch.empty(bsk_bytes // 8, dtype=torch.float64, device=device)
    ksk_buf = torch.empty(ksk_bytes // 8, dtype=torch.int64, device=device)
    if rank == 0:
        bsk_buf.copy_(device_view(bsk.device_ptr, bsk_bytes // 8, "<f8", device))
        ksk_buf.copy_(device_view(ksk.device_ptr, ksk_bytes // 8, "<i8", device))
        torch.cuda.synchronize(device)
    broadcast_buffer(bsk_buf, 0)
    broadcast_buffer(ksk_buf, 0)
    if rank == 0:
        return bsk, ksk
    return api.BootstrapKey.adopt(params, bsk_buf), api.KeySwitchKey.adopt(params, ksk_buf)


def device_view(ptr: int, count: int, typestr: str, device):
    """torch tensor over library-owned device memory (no copy)."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device=device)


def gather_results(local: np.ndarray, count: int) -> np.ndarray | None:
    """Rank 0 receives every rank's output rows in shard order (host side; used by the e2e path)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = shard_bounds(count, world)
    width = local.shape[1]
    pad = max(e - b for b, e in bounds)
    buf = np.zeros((pad, width), np.int64)
    buf[: local.shape[0]] = local.view(np.int64)
    t = torch.from_numpy(buf)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    outs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, outs, dst=0)
    if rank != 0:
        return None
    return np.concatenate([o.cpu().numpy()[: e - b] for o, (b, e) in zip(outs, bounds)]).view(np.uint64)
