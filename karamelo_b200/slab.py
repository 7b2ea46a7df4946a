"""One process per GPU: launcher-side plumbing for the slab-decomposed engine.

torch.distributed is used only for process bootstrap (rendezvous, broadcasting the NCCL unique
id, gathering results for tests and the max-over-ranks timing); the data path - halo sums of the
shared node planes, particle migration, the dt reduction - is NCCL inside libkml.so
(karamelo_b200/csrc/kml_comm.cuh).
"""
import os

import numpy as np


def init_distributed():
    """(rank, world, local_rank, dist) from the torchrun environment; nccl on GPUs, gloo on CPU."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        use_cuda = torch.cuda.is_available()
        if use_cuda:
            torch.cuda.set_device(local)
        dist.init_process_group("nccl" if use_cuda else "gloo", rank=rank, world_size=world)
    return rank, world, local, dist


def make_engine(lib=None, quiet=True, with_comm=True):
    """Engine for this rank's slab; rank 0 creates the NCCL unique id and broadcasts it."""
    import torch
    from .api import Engine, load_host_library
    rank, world, local, dist = init_distributed()
    if world == 1:
        return Engine(lib, quiet=quiet, device=local)
    lib = lib if lib is not None and not isinstance(lib, str) else load_host_library(lib)
    nccl_id = None
    if with_comm:
        import ctypes as C
        buf = C.create_string_buffer(128)
        if rank == 0 and lib.kml_comm_unique_id(C.cast(buf, C.c_void_p)):
            raise RuntimeError(lib.kml_last_error().decode())
        t = torch.tensor(list(buf.raw), dtype=torch.uint8)
        if torch.cuda.is_available():
            t = t.cuda()
        dist.broadcast(t, src=0)
        nccl_id = bytes(t.cpu().tolist())
    return Engine(lib, quiet=quiet or rank != 0, device=local, rank=rank, nranks=world, nccl_id=nccl_id)


def gather_snapshot(engine, fields, solid=0):
    """All ranks' particles of one solid concatenated and sorted by tag, on every rank."""
    from .api import P
    _, world, _, dist = init_distributed()
    d = {f: engine.download(solid, getattr(P, f)) for f in fields}
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, d)
        d = {f: np.concatenate([p[f] for p in parts]) for f in fields}
    order = np.argsort(d["PTAG"], kind="stable")
    return {k: v[order] for k, v in d.items()}
