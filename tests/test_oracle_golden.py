"""The CPU oracle (oracle/oracle_kml.cpp) against golden state written by the UNMODIFIED
reference (tests/golden/*.npz, made by tests/golden/make_golden.py from oracle/_ref).

This is what pins the oracle: every BASELINE config variant must reproduce the reference's
particle state after 100 steps.  UL cases are bit-exact; TL cases differ only through the
Eigen stand-ins (JacobiSVD / EigenSolver are an un-vendored third-party dependency), so they
get a 1e-12 relative tolerance.
"""
import pytest

from cases import CASES
from common import compare_to_golden, load_golden, run_case

TOL = {name: (1e-12 if CASES[name][1] else 0.0) for name in CASES}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference(oracle_lib, name):
    script, is_tl, thermal, steps = CASES[name]
    golden, log_last = load_golden(name)
    snap, st = run_case(oracle_lib, script, steps, thermal)
    assert st["ntimestep"] == steps
    worst = compare_to_golden(snap, golden, TOL[name] if TOL[name] > 0 else 1e-300)
    # dt / time of the reference log (6 significant digits)
    assert abs(st["dt"] - log_last[1]) <= 1e-5 * abs(log_last[1])
    print(name, worst)
