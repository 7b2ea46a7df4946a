"""The reference's neighbour-membership test and what it does to reproducibility.

ULMPM::compute_grid_weight_functions_and_gradients keeps node I in particle p's neighbour list only `if (wf != 0)`
(src/ulmpm.cpp:252-263).  For cubic B-splines the weight of the outermost stencil node is (2 - r)^3 / 6 evaluated in Horner form
(src/basis_functions.h): for a particle within ~8e-6 h of leaving that node the true weight is below 1e-16 and the computed one is 0
or a multiple of 2.2e-16 depending on the last bit of r - while the gradient, (2 - r)^2 / 2h ~ 2e-11 / h, is far from zero.  Whenever
the weight rounds to exactly 0 the reference drops the node and with it |v| dw dt ~ 3e-13 of the particle's deformation gradient,
which the stress shows as ~1e-10 of max|sigma| (bulk modulus over yield stress).  Round 1 measured exactly this on 4 slabs (1.03e-10)
and 8 slabs (1.72e-10): the cell kernels evaluated all 64 stencil nodes unconditionally.  They now mirror the test (dw = 0 where
w == 0, kml_p2g_cell3.cuh cubic_axis4).  What stays is the reference's own coin flip: one ulp of the particle position decides the
branch, so two runs of the REFERENCE algorithm that differ only in summation order disagree by the same amount (second test).
"""
import numpy as np
import pytest

from cases import block
from common import oracle_both_memberships, permute_particles, rel, rel_either
from karamelo_b200.api import Engine

FIELDS = ("PTAG", "X", "V", "SIGMA", "FDEF", "EFF_PLASTIC_STRAIN", "VOL")


def test_membership_test_explains_the_slab_disagreement(oracle_lib):
    """24 x 6 x 6 cells, squeeze rate 2.5e-3, drift 0.03 (the four-slab case of test_slab.py): dropping vs keeping zero-weight nodes moves
    the stress by 1.03e-10 and F by 3.3e-13 - the numbers the four-slab CUDA run showed against the oracle in round 1."""
    (ref, _), (keep, _) = oracle_both_memberships(oracle_lib, block((24, 6, 6), "musl", "cubic-spline", a=2.5e-3, drift=0.03), 100, FIELDS)
    d = {k: rel(keep[k], ref[k]) for k in FIELDS if k != "PTAG"}
    print(d)
    assert 0.9e-10 < d["SIGMA"] < 1.2e-10 and 2e-13 < d["FDEF"] < 5e-13
    # one particle carries it
    per_particle = np.abs(keep["FDEF"] - ref["FDEF"]).reshape(len(ref["PTAG"]), -1).max(1)
    print("particles with |dF| > 1e-13:", int((per_particle > 1e-13).sum()), "of", len(per_particle))
    assert (per_particle > 1e-13).sum() <= 16  # the particles of the cell(s) that took the other branch


def test_reference_algorithm_flips_with_summation_order(oracle_lib):
    """48 x 6 x 6 cells at the benchmark's rate (the eight-slab case): the oracle against itself with the particle arrays shuffled -
    same arithmetic per (particle, node), different summation order - differs by 1.5e-10 in the stress, because a particle position
    that differs in the last bit takes the other branch of `wf != 0`.  With zero-weight nodes kept the two runs agree to 1e-12."""
    script = block((48, 6, 6), "musl", "cubic-spline", a=2.5e-4, drift=0.03)

    def run(seed):
        e = Engine(oracle_lib)
        e.script(script)
        if seed:
            permute_particles(e, seed)
        e.line("run(100)")
        s = e.snapshot(FIELDS)[0]
        e.close()
        return s
    a, b = run(0), run(1)
    d = {k: rel(b[k], a[k]) for k in FIELDS if k != "PTAG"}
    print("reference semantics:", d)
    assert d["SIGMA"] > 2e-11, "the case no longer contains a marginal membership event - pick another one"
    import os
    os.environ["KML_ORACLE_KEEP_ZERO_WEIGHT"] = "1"
    try:
        a, b = run(0), run(1)
    finally:
        os.environ.pop("KML_ORACLE_KEEP_ZERO_WEIGHT", None)
    d = {k: rel(b[k], a[k]) for k in FIELDS if k != "PTAG"}
    print("zero-weight nodes kept:", d)
    assert max(d.values()) < 5e-12


@pytest.mark.gpu
@pytest.mark.parametrize("cells,a", [((24, 6, 6), 2.5e-3), ((48, 6, 6), 2.5e-4)])
def test_cuda_follows_one_of_the_two_branches(cuda_lib, oracle_lib, cells, a):
    """One GPU, the two cases above: every element of the CUDA result is within 1e-10 of the oracle with the membership test or of the
    oracle without it (an element-wise minimum, not a looser tolerance)."""
    script = block(cells, "musl", "cubic-spline", a=a, drift=0.03)
    (ref, st_ref), (keep, _) = oracle_both_memberships(oracle_lib, script, 100, FIELDS)
    e = Engine(cuda_lib)
    e.script(script + "\nrun(100)\n")
    got = e.snapshot(FIELDS)[0]
    st = e.state()
    e.close()
    assert (got["PTAG"] == ref["PTAG"]).all()
    d = {k: rel_either(got[k], ref[k], keep[k]) for k in FIELDS if k != "PTAG"}
    print(cells, a, "either-branch error", d, "vs reference branch only", {k: rel(got[k], ref[k]) for k in FIELDS if k != "PTAG"})
    # the CUDA run can take a flip of its own that neither oracle run has (its particle positions differ from the oracle's in the last bit):
    # the stress is held to 1e-10 plus the bound of one flip (tests/slab_worker.py), everything else to 1e-10
    assert max(v for k, v in d.items() if k != "SIGMA") <= 1e-10 and d["SIGMA"] <= 7e-10
    assert abs(st["dt"] - st_ref["dt"]) <= 1e-10 * st_ref["dt"]
