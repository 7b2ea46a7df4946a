"""Diagnostic (not a test): which cell kernel disagrees with the oracle on a case, and where.
    python tests/diag_parity.py <case> [steps]     (GPU box)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cases import CASES  # noqa: E402
from common import rel, run_case  # noqa: E402
from conftest import ORACLE_HOST_LIB  # noqa: E402
from karamelo_b200.api import load_host_library  # noqa: E402

name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
script = CASES[name][0]
ref, _ = run_case(load_host_library(ORACLE_HOST_LIB), script, steps)
for mask in (0, 1, 2, 4, 7):
    os.environ["KML_CELL_MASK"] = str(mask)
    got, _ = run_case(load_host_library(None), script, steps)
    w = {k: rel(got[0][k], ref[0][k]) for k in ("X", "V", "SIGMA", "FDEF")}
    d = np.abs(got[0]["V"] - ref[0]["V"]).max(1)
    bad = np.argsort(d)[-3:]
    print("mask", mask, {k: "%.1e" % v for k, v in w.items()}, "worst particles (x0):", [tuple(np.round(ref[0]["X"][i], 2)) for i in bad])
