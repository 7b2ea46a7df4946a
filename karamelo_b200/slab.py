"""One process per GPU: launcher-side plumbing for the slab-decomposed engine.

torch.distributed is used only for process bootstrap (rendezvous, broadcasting the NCCL unique
id, gathering results for tests and the max-over-ranks timing); the data path - halo sums of the
shared node planes, particle migration, the dt reduction - is NCCL inside libkml.so
(karamelo_b200/csrc/kml_comm.cuh).
"""
import json
import os
import time

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


def gather_snapshot(engine, fields):
    """All ranks' particles (solid 0) concatenated and sorted by tag, on every rank."""
    from .api import P
    _, world, _, dist = init_distributed()
    d = {f: engine.download(0, getattr(P, f)) for f in fields}
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, d)
        d = {f: np.concatenate([p[f] for p in parts]) for f in fields}
    order = np.argsort(d["PTAG"], kind="stable")
    return {k: v[order] for k, v in d.items()}


def bench_multi_gpu(args):
    """bench.py at N > 1: the same 100M-particle block, slab-decomposed (strong scaling).  Timing: barrier +
    device sync on both sides, max over ranks; rank 0 prints the JSON line."""
    import torch
    from . import api
    import bench as B
    rank, world, local, dist = init_distributed()
    cells = tuple(args.cells)
    W, K = max(args.warmup, 3), args.steps
    eng = make_engine(None)
    t0 = time.perf_counter()
    eng.script(B.block_script(cells, velocity_fix=False))
    x = eng.download(0, api.P.X)
    eng.upload(0, api.P.V, B.squeeze_velocity(x, cells))
    np_local = len(x)
    del x
    setup_s = time.perf_counter() - t0
    npart = eng.slab_info(0)["np_global"]
    eng.line("run(%d)" % W)
    eng.stage_times(reset=True)
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    dist.barrier()
    eng.timer_start()
    eng.line("run(%d)" % K)
    ms_local = eng.timer_stop()
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    counts = eng.stage_times(reset=True)
    launches = int(sum(v[1] for v in counts.values()))
    eng.profile(True)
    eng.line("run(3)")
    st = eng.stage_times(reset=True)
    eng.profile(False)
    nps = [None] * world
    dist.all_gather_object(nps, (np_local, {k: v[0] / 3 for k, v in st.items()}))
    flags = eng.error_flags()
    if rank == 0:
        peak, peak_src = B.measured_peak()
        stage_ms = {k: max(p[1][k] for p in nps) for k in nps[0][1]}
        per_stage = {k: {"ms": round(stage_ms[k], 4), "algo_GBps": round(b * npart / world / (stage_ms[k] * 1e-3) / 1e9, 1)} for k, b in B.ALGO_BYTES.items() if stage_ms.get(k, 0) > 0}
        dom = max(per_stage, key=lambda k: per_stage[k]["ms"])
        line = {"metric": "particle_steps_per_sec", "value": npart * K / (ms * 1e-3), "unit": "particle-steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "synthetic 3-D ULMPM elastoplastic block (configs[4]): cubic B-splines, MUSL, FLIP 0.99, linear EOS + plastic strength, adaptive dt",
                           "cells": list(cells), "particles": npart, "particles_per_rank": [p[0] for p in nps], "parallelism": "x-slab x%d, NCCL halo sums + migration" % world,
                           "l2": "inputs >> L2 (no flush needed)", "setup_s": round(setup_s, 1)},
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": per_stage[dom]["algo_GBps"], "peak": peak, "unit": "GB/s", "frac": round(per_stage[dom]["algo_GBps"] / peak, 4),
                             "traffic": None, "peak_source": peak_src, "per_stage_max_over_ranks": per_stage, "note": "per-GPU algorithmic GB/s"},
                "cpu_baseline": None, "e2e": None, "clocks": sampler.summary(), "gpu_launches": launches, "error_flags": flags}
        print(json.dumps(line))
    eng.close()
    dist.barrier()
