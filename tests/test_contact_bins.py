"""Cell-list contact (kml_kernels.cuh k_contact_bins): solid 2 binned by the reference's first screen (|dx_i| < max_cellsize,
src/fix_contact_hertz.cpp:117-127, src/fix_contact_min_penetration.cpp:124-134), every particle of solid 1 visits the 3^dim bins around
its own.  Same screens and pair set as the all-pairs sweep (k_contact), which the engine keeps for small bodies; KML_CONTACT=bins | pairs
forces a path."""
import time

import pytest

from cases import CASES, bouncing_balls
from common import FIELDS, compare_snaps, compare_to_golden, load_golden, run_case
from karamelo_b200.api import Engine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c4_balls_hertz", "c4_balls_minpen", "x_rigid_tl_contact"])
def test_binned_contact_matches_the_reference(cuda_lib, name, monkeypatch):
    monkeypatch.setenv("KML_CONTACT", "bins")
    script, is_tl, thermal, steps = CASES[name]
    got, _ = run_case(cuda_lib, script, steps, thermal)
    golden, _ = load_golden(name)
    print(name, compare_to_golden(got, golden, 1e-10))


@pytest.mark.parametrize("contact", ["hertz", "minimize_penetration"])
def test_binned_contact_on_large_bodies(cuda_lib, contact, monkeypatch):
    """Two discs of 51 000 particles each (N = 320: 2.6e9 pair tests per step for the all-pairs sweep): both paths agree, and the timing of each
    is printed for the record."""
    script = bouncing_balls(contact).replace("N        = 40", "N        = 320").replace("set_dt(0.001)", "set_dt(0.000125)").replace("c = 0.16", "c = 0.142")  # the discs touch after ~15 steps
    out = {}
    for mode in ("pairs", "bins"):
        monkeypatch.setenv("KML_CONTACT", mode)
        e = Engine(cuda_lib)
        e.script(script + "\nrun(2)\n")
        e.synchronize()
        t0 = time.perf_counter()
        e.line("run(40)")
        e.synchronize()
        dt = (time.perf_counter() - t0) / 40
        out[mode] = (e.snapshot(FIELDS), dt, sum(e.solid_info(i)["np"] for i in range(e.nsolids())))
        e.close()
    print(contact, "particles", out["bins"][2], "ms/step: all pairs %.3f, bins %.3f" % (out["pairs"][1] * 1e3, out["bins"][1] * 1e3),
          compare_snaps(out["bins"][0], out["pairs"][0], 1e-10))
    import numpy as np
    assert max(float(np.abs(s["SIGMA"]).max()) for s in out["bins"][0]) > 0, "the discs never touched: the comparison would be vacuous"
    assert out["bins"][1] < out["pairs"][1]
