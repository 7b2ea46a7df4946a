#!/usr/bin/env python
"""bench.py - particle-steps/s of the full MPM step on the synthetic 3-D ULMPM elastoplastic block
(BASELINE.json configs[4]: cubic B-splines, MUSL, 8 particles per cell, 248x250x202 cells = 100 192 000
particles), through the host driver + C ABI of include/kml.h.

    python bench.py --gpus N --steps K --warmup W            our arm (CUDA engine)
    python bench.py --impl reference ...                      the reference's own CPU path (oracle/_ref)

Prints ONE JSON line (see the keys at the bottom).  A "step" is one full MUSL step
(re-bin, P2G, grid update, G2P+advance, MUSL re-projection, gradient+F+stress, dt reduction).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FULL_CELLS = (248, 250, 202)  # 100 192 000 particles in a 256 x 258 x 210 box
# algorithmic bytes per particle-step (fp64 SoA, every array once per kernel)
ALGO_BYTES = {"rebin": 80, "p2g": 158, "grid": 18, "g2p": 103, "v2g": 64, "stress": 436}
A_SQUEEZE = 2.5e-4


def block_script(cells, velocity_fix):
    from cases import block
    s = block(tuple(cells), "musl", "cubic-spline", fixed_dt=False, a=A_SQUEEZE)
    if not velocity_fix:
        s = "\n".join(ln for ln in s.splitlines() if not ln.startswith("fix(v0")) + "\n"
    return s


def squeeze_velocity(x, cells):
    """isochoric uniaxial squeeze, linear in position; a goes through the script's float literal path"""
    a = float(np.float32(2.5)) * 10.0 ** -4
    c = np.array([4 + cells[0] / 2, 4 + cells[1] / 2, 4 + cells[2] / 2])
    v = np.empty_like(x)
    v[:, 0] = -a * (x[:, 0] - c[0])
    v[:, 1] = 0.5 * a * (x[:, 1] - c[1])
    v[:, 2] = 0.5 * a * (x[:, 2] - c[2])
    return v


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm)}


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(sample_cells=(24, 24, 24), nsteps=6):
    """The reference's CPU path on a bounded sample of the same workload, on this box's host cores.
    kind 'reference' = the unmodified reference binary (oracle/_ref), else 'port' = oracle/oracle_kml.cpp."""
    script = block_script(sample_cells, velocity_fix=True)
    npart = sample_cells[0] * sample_cells[1] * sample_cells[2] * 8
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "karamelo_ref")
    import tempfile
    if os.path.exists(ref_bin):
        def run(n):
            d = tempfile.mkdtemp(prefix="kmlbench_")
            open(os.path.join(d, "in.mpm"), "w").write(script + "run(%d)\n" % n)
            t0 = time.perf_counter()
            subprocess.run([ref_bin, "-i", "in.mpm"], cwd=d, capture_output=True, check=True)
            return time.perf_counter() - t0
        t1 = run(1)
        t2 = run(1 + nsteps)
        per_step = max(t2 - t1, 1e-9) / nsteps
        kind = "reference"
    else:
        from karamelo_b200.api import Engine
        e = Engine(os.path.join(ROOT, "oracle", "_build", "libkml_host_oracle.so"))
        e.script(script + "run(1)\n")
        t0 = time.perf_counter()
        e.line("run(%d)" % nsteps)
        per_step = (time.perf_counter() - t0) / nsteps
        e.close()
        kind = "port"
    return {"value": npart / per_step, "unit": "particle-steps/s", "cores": 1, "kind": kind,
            "sample": "same script at %dx%dx%d cells (%d particles), %d MUSL steps, 1 MPI-less rank" % (sample_cells + (npart, nsteps))}


def run_reference_arm(args):
    cb = cpu_baseline(tuple(args.ref_cells), max(args.steps, 1))
    line = {"impl": "reference", "metric": "particle_steps_per_sec", "value": cb["value"], "unit": "particle-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic 3-D ULMPM elastoplastic block, cubic B-splines, MUSL, 8 ppc (bounded sample for the CPU)", "sample": cb["sample"]},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, nargs=3, default=list(FULL_CELLS))
    ap.add_argument("--ref-cells", type=int, nargs=3, default=[24, 24, 24])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            run_reference_arm(args)
        return

    if world > 1 or args.gpus > 1:
        from karamelo_b200.slab import bench_multi_gpu
        bench_multi_gpu(args)
        return

    import torch
    from karamelo_b200.api import Engine, P
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the engine has no CPU fallback"
    cells = tuple(args.cells)
    W, K = max(args.warmup, 3), args.steps
    eng = Engine(device=0)
    t0 = time.perf_counter()
    eng.script(block_script(cells, velocity_fix=False))
    npart = eng.solid_info(0)["np"]
    x = eng.download(0, P.X)
    eng.upload(0, P.V, squeeze_velocity(x, cells))
    del x
    setup_s = time.perf_counter() - t0

    eng.line("run(%d)" % W)           # warm-up (>= 3 steps); includes step 1 with dt = 1e-16 like the reference
    eng.stage_times(reset=True)
    sampler = ClockSampler(0)
    sampler.start()
    eng.synchronize()
    eng.timer_start()                 # CUDA events on the engine's stream
    eng.line("run(%d)" % K)           # K full steps; state (~45 GB at 100M particles) is far larger than the 126 MB L2
    ms = eng.timer_stop()
    clocks = sampler.summary()
    counts = eng.stage_times(reset=True)
    launches = int(sum(v[1] for v in counts.values()))
    value = npart * K / (ms * 1e-3)

    # per-stage device time (events around every stage, a few extra steps) -> roofline of the dominant kernel
    eng.profile(True)
    eng.line("run(3)")
    st = eng.stage_times(reset=True)
    eng.profile(False)
    stage_ms = {k: v[0] / 3 for k, v in st.items()}
    peak, peak_src = measured_peak()
    per_stage = {}
    for k, b in ALGO_BYTES.items():
        if stage_ms.get(k, 0) > 0:
            per_stage[k] = {"ms": round(stage_ms[k], 4), "algo_GBps": round(b * npart / (stage_ms[k] * 1e-3) / 1e9, 1), "frac": round(b * npart / (stage_ms[k] * 1e-3) / 1e9 / peak, 4)}
    dom = max((k for k in ALGO_BYTES if stage_ms.get(k, 0) > 0), key=lambda k: stage_ms[k])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": per_stage[dom]["algo_GBps"], "peak": peak, "unit": "GB/s", "frac": per_stage[dom]["frac"],
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_particle": ALGO_BYTES[dom], "per_stage": per_stage}

    # end to end through the public API with HOST buffers: every step uploads the step's particle inputs from pinned
    # host memory, runs one step, downloads the results
    e2e = None
    if not args.no_e2e:
        fields_in = [P.X, P.V, P.SIGMA, P.FDEF, P.VOL, P.EFF_PLASTIC_STRAIN, P.EFF_PLASTIC_STRAIN_RATE]
        fields_out = [P.X, P.V, P.SIGMA, P.EFF_PLASTIC_STRAIN]
        host = {}
        for f in fields_in:
            a = eng.download(0, f)
            t = torch.empty(a.shape, dtype=torch.float64, pin_memory=True)
            t.numpy()[...] = a
            host[f[0]] = t
        bi = sum(host[f[0]].numel() * 8 for f in fields_in)
        bo = sum(host[f[0]].numel() * 8 for f in fields_out)
        import ctypes as C
        info = eng.solid_info(0)
        eng.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            for f in fields_in:
                eng._ckk(eng.lib.kml_solid_upload(eng.ctx, info["solid"], f[0], C.c_void_p(host[f[0]].data_ptr())))
            eng.line("run(1)")
            for f in fields_out:
                eng._ckk(eng.lib.kml_solid_download(eng.ctx, info["solid"], f[0], C.c_void_p(host[f[0]].data_ptr())))
        eng.synchronize()
        dt_e2e = time.perf_counter() - t0
        e2e = {"value": npart * args.e2e_steps / dt_e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo, "steps": args.e2e_steps}
        del host

    flags = eng.error_flags()
    eng.close()
    cb = None if args.no_cpu_baseline else cpu_baseline(tuple(args.ref_cells))
    line = {"metric": "particle_steps_per_sec", "value": value, "unit": "particle-steps/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic 3-D ULMPM elastoplastic block (configs[4]): cubic B-splines, MUSL, FLIP 0.99, linear EOS + plastic strength, adaptive dt",
                       "cells": list(cells), "particles": npart, "particles_per_cell": 8, "l2": "inputs >> L2 (no flush needed)", "setup_s": round(setup_s, 1)},
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "clocks": clocks, "gpu_launches": launches, "error_flags": flags}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
