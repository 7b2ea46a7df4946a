#!/usr/bin/env python
"""Golden fixture for the headline block at SURVEY section 8d's parity size (32^3 cells, 262 144 particles,
100 MUSL steps, adaptive dt), from the UNMODIFIED reference (oracle/_ref/karamelo_ref).

The full state is 55 MB; the fixture keeps every 64th particle by tag (4 096 rows: tag, x, v, sigma, F, eps,
epsdot) at steps 20 and 100, plus whole-population sums as a "checksum of checksums" (compared with a tolerance
that allows for the summation order).  Needs /root/reference (build container only).

    python tests/golden/make_golden_large.py
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cases import block  # noqa: E402
from refio import REF_BIN, read_restart_solids  # noqa: E402

CELLS, STEPS, EVERY, STRIDE = (32, 32, 32), (20, 100), 20, 64
SYM = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]


def script():
    return block(CELLS, "musl", "cubic-spline", fixed_dt=False, a=2.5e-3)


def reduce_state(s):
    """subsample + sums of one solid's restart record (shared with the tests through the same keys)"""
    order = np.argsort(s["ptag"], kind="stable")
    keep = order[(s["ptag"][order] - 1) % STRIDE == 0]
    sig6 = np.stack([s["sigma"][:, a, b] for a, b in SYM], axis=1)
    out = {"ptag": s["ptag"][keep], "x": s["x"][keep], "v": s["v"][keep], "sigma": sig6[keep], "F": s["F"].reshape(-1, 9)[keep],
           "eps": s["eps"][keep], "epsdot": s["epsdot"][keep]}
    out["sum_x"] = s["x"].sum(0)
    out["sum_absv"] = np.abs(s["v"]).sum(0)
    out["sum_abssigma"] = np.abs(sig6).sum(0)
    out["sum_eps"] = np.array([s["eps"].sum()])
    out["n_plastic"] = np.array([int((s["eps"] > 0).sum())])
    out["np"] = np.array([len(s["ptag"])])
    return out


def main():
    d = tempfile.mkdtemp(prefix="kmlref_large_")
    open(os.path.join(d, "in.mpm"), "w").write(script() + "\nrestart(%d, ref-*.restart)\nrun(%d)\n" % (EVERY, max(STEPS)))
    p = subprocess.run([REF_BIN, "-i", "in.mpm"], cwd=d, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(p.stdout[-3000:] + p.stderr[-2000:])
    nps = [int(m) for m in re.findall(r"^np_local=(\d+)", p.stdout, flags=re.M)]
    data = {"cells": np.array(CELLS), "stride": np.array([STRIDE])}
    for st in STEPS:
        s = read_restart_solids(os.path.join(d, "ref-%d.restart" % st), nps, False, False)[0]
        for k, v in reduce_state(s).items():
            data["%s_%d" % (k, st)] = v
        print("step %d: np=%d plastic=%d max eps=%.3g" % (st, len(s["ptag"]), int((s["eps"] > 0).sum()), s["eps"].max()))
    path = os.path.join(HERE, "large_c5_block_32.npz")
    np.savez_compressed(path, **data)
    print("->", path, os.path.getsize(path) // 1024, "kB")


if __name__ == "__main__":
    main()
