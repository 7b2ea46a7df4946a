"""Parity of the CUDA engine (through the C ABI + host driver) with the CPU oracle and with
the golden state of the unmodified reference, after 100 steps in fp64.

Bar (BASELINE.json north_star): particle counts / tags bit-exact; positions, velocities,
stresses and plastic strain within 1e-10 relative (relative to the field magnitude) - not
bit-equality because the atomic / re-ordered node sums change the summation order.
"""
import numpy as np
import pytest

from cases import CASES
from common import compare_snaps, compare_to_golden, load_golden, run_case

TOL = 1e-10
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_matches_oracle_and_reference(cuda_lib, oracle_lib, name):
    assert cuda_lib.kml_backend().decode().startswith("cuda"), "the product library must be the CUDA engine"
    script, is_tl, thermal, steps = CASES[name]
    got, st = run_case(cuda_lib, script, steps, thermal)
    ref, st_ref = run_case(oracle_lib, script, steps, thermal)
    assert st["ntimestep"] == st_ref["ntimestep"] == steps
    worst = compare_snaps(got, ref, TOL)
    assert abs(st["dt"] - st_ref["dt"]) <= TOL * abs(st_ref["dt"])
    assert abs(st["time"] - st_ref["time"]) <= TOL * abs(st_ref["time"])
    golden, _ = load_golden(name)
    worst_g = compare_to_golden(got, golden, TOL)
    print(name, "vs oracle", worst, "vs reference", worst_g)


def test_indexing_bit_exact(cuda_lib, oracle_lib):
    """Node types, node positions and particle lattice are integer/bit-exact with the oracle."""
    from karamelo_b200.api import Engine, N, P
    from cases import taylor_bar
    a, b = Engine(cuda_lib), Engine(oracle_lib)
    for e in (a, b):
        e.script(taylor_bar("cubic-spline"))
    assert a.solid_info(0)["n"] == b.solid_info(0)["n"]
    assert (a.grid_download(0, N.NTYPE) == b.grid_download(0, N.NTYPE)).all()
    assert (a.grid_download(0, N.X0) == b.grid_download(0, N.X0)).all()
    assert (a.grid_download(0, N.MASK) == b.grid_download(0, N.MASK)).all()
    for f in (P.PTAG, P.X, P.X0, P.MASS, P.VOL0, P.MASK):
        assert (a.download(0, f) == b.download(0, f)).all()
