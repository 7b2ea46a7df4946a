"""bench.py's end-to-end protocol (e2e_loop: per step upload the particle inputs from host buffers, run one step, read the results back)
must not change the physics: the host buffers always hold the engine's own state, so a run interleaved with those round trips is
bit-identical to an uninterrupted one.  Run here on the CPU oracle through the same C ABI entry points (kml_solid_upload / _download /
_generation); on the GPU the same loop is the timed e2e region of bench.py."""
import os
import sys

import numpy as np

from cases import block
from common import FIELDS
from conftest import ROOT
from karamelo_b200.api import Engine

sys.path.insert(0, ROOT)


def test_e2e_round_trips_do_not_change_the_run(oracle_lib):
    import bench
    script = block((6, 5, 4), "musl", "cubic-spline", a=2.5e-3)
    a = Engine(oracle_lib)
    a.script(script + "run(5)\n")
    h2d, d2h = bench.e2e_loop(a, 4, pinned=False)
    b = Engine(oracle_lib)
    b.script(script + "run(9)\n")
    sa, sb = a.snapshot(FIELDS)[0], b.snapshot(FIELDS)[0]
    for k in FIELDS:
        assert (sa[k] == sb[k]).all(), k
    n = len(sa["PTAG"])
    assert h2d == n * 27 * 8 and d2h == n * 27 * 8  # 27 doubles per particle each way
    assert a.state()["ntimestep"] == b.state()["ntimestep"] == 9
    a.close()
    b.close()
