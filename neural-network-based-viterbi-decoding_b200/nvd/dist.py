"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous, and the single
exchange step of the path — an all-reduce(sum) of the [n_snr, 4] error counters (ncclAllReduce
over NVLink with the `nccl` backend; `gloo` on CPU for the world_size-2 tests).  Frames are
sharded by global index range (montecarlo.shard_range), so there is no data-path collective."""
from __future__ import annotations

import os

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank)."""
    rank, world, local = env_rank_world()
    if world == 1:
        return rank, world, local
    import torch
    import torch.distributed as dist

    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def all_reduce_sum(arr: np.ndarray) -> np.ndarray:
    """Sum an int64 numpy array over all ranks (identity when not initialised)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return arr
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def all_reduce_max(value: float) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def broadcast_bytes(data: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Ship a small byte string from rank `src` to every rank (the NCCL unique id of nvd_comm_init)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return bytes(data)
    t = torch.zeros(nbytes, dtype=torch.uint8)
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(data), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def all_gather_bytes(data: bytes, nbytes: int) -> bytes:
    """Every rank contributes `nbytes` bytes; returns the concatenation in rank order (the CUDA IPC handles of
    nvd_comm_p2p_open)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return bytes(data)
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8).clone()
    out = [torch.zeros(nbytes, dtype=torch.uint8) for _ in range(dist.get_world_size())]
    if dist.get_backend() == "nccl":
        t = t.cuda()
        out = [o.cuda() for o in out]
    dist.all_gather(out, t)
    return b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)


def barrier():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shutdown():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()
