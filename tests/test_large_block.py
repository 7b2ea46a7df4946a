"""The headline block at SURVEY section 8d's parity size: 32^3 cells = 262 144 particles, cubic B-splines, MUSL,
adaptive dt, 100 steps, against golden state of the UNMODIFIED reference (tests/golden/large_c5_block_32.npz:
every 64th particle by tag at steps 20 and 100 plus whole-population sums; tests/golden/make_golden_large.py).

CPU: the oracle after 20 steps (bit-exact on the sample).  GPU: the CUDA engine after 20 and 100 steps (tags exact,
fields within 1e-10 of the field magnitude; the sums within 1e-9 - their own summation order differs)."""
import os

import numpy as np
import pytest

from common import GOLDEN_DIR, SYM, rel
from karamelo_b200.api import Engine

import sys
sys.path.insert(0, GOLDEN_DIR)
from make_golden_large import STRIDE, script  # noqa: E402

FIELDS = ("PTAG", "X", "V", "SIGMA", "FDEF", "EFF_PLASTIC_STRAIN", "EFF_PLASTIC_STRAIN_RATE")


def _golden():
    return np.load(os.path.join(GOLDEN_DIR, "large_c5_block_32.npz"))


def _check(snap, g, step, tol):
    s = snap[0]
    assert len(s["PTAG"]) == int(g["np_%d" % step][0]) == 262144
    keep = (s["PTAG"] - 1) % STRIDE == 0
    assert (s["PTAG"][keep] == g["ptag_%d" % step]).all(), "tags of the sample differ"
    sig6 = np.stack([s["SIGMA"][:, a, b] for a, b in SYM], 1)
    pairs = {"x": s["X"][keep], "v": s["V"][keep], "sigma": sig6[keep], "F": s["FDEF"].reshape(-1, 9)[keep],
             "eps": s["EFF_PLASTIC_STRAIN"][keep], "epsdot": s["EFF_PLASTIC_STRAIN_RATE"][keep]}
    worst = {k: rel(v, g["%s_%d" % (k, step)]) for k, v in pairs.items()}
    bad = {k: v for k, v in worst.items() if v > tol}
    assert not bad, (step, bad, worst)
    # whole-population checks: every particle contributes
    assert int((s["EFF_PLASTIC_STRAIN"] > 0).sum()) == int(g["n_plastic_%d" % step][0])
    sums = {"sum_x": s["X"].sum(0), "sum_absv": np.abs(s["V"]).sum(0), "sum_abssigma": np.abs(sig6).sum(0),
            "sum_eps": np.array([s["EFF_PLASTIC_STRAIN"].sum()])}
    for k, v in sums.items():
        assert rel(v, g["%s_%d" % (k, step)]) <= max(tol, 1e-9), (step, k, v, g["%s_%d" % (k, step)])
    return worst


def test_oracle_matches_reference_at_parity_size(oracle_lib):
    g = _golden()
    e = Engine(oracle_lib)
    e.script(script() + "\nrun(20)\n")
    worst = _check(e.snapshot(FIELDS), g, 20, 1e-300)
    e.close()
    print(worst)


@pytest.mark.gpu
def test_cuda_matches_reference_at_parity_size(cuda_lib):
    g = _golden()
    e = Engine(cuda_lib)
    e.script(script() + "\nrun(20)\n")
    w20 = _check(e.snapshot(FIELDS), g, 20, 1e-10)
    e.line("run(80)")
    w100 = _check(e.snapshot(FIELDS), g, 100, 1e-10)
    assert e.error_flags() == 0
    e.close()
    print("step 20", w20, "step 100", w100)
