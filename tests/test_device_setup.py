"""Set-up on the device: the particle lattice of Solid::populate (src/solid.cpp:1810-2336), particle group
masks (Group::assign, src/group.cpp:65-238) and initial_velocity_particles expressions (src/fix_initial_velocity_particles.cpp:99-161)
evaluated by kernels must be BIT-identical to the host driver's restated reference loops (KML_HOST_POPULATE=1 selects those), which
tests/test_oracle_golden.py pins against the unmodified reference."""
import os

import numpy as np
import pytest

from cases import CASES, axisym_bar, block, taylor_bar, two_disks, two_spheres
from karamelo_b200.api import Engine, P

FIELDS = ("PTAG", "X", "X0", "V", "MASS", "VOL0", "VOL", "RHO0", "MASK", "FDEF", "T")
SCRIPTS = {
    "block_32": block((32, 32, 32), "musl", a=2.5e-3),
    "block_drift_ppc3": block((5, 6, 7), "usl", ppc=3, drift=0.03),
    "block_ppc1_margin1": block((6, 5, 4), "musl", ppc=1, margin=1),
    "taylor_bar_cylinder_x": taylor_bar("cubic-spline"),
    "two_disks_2d": two_disks("musl"),
    "two_spheres": two_spheres(),
    "axisymmetric": axisym_bar(),
    "tensile_tl_thermal": CASES["p_thermal_full_tl"][0],
    "exterior_region": two_disks("musl").replace("region(rBall2, cylinder,  c,  c, R)", "region(rBall2, cylinder,  c,  c, R, exterior)").replace(
        "solid(sBall2, region, rBall2", "region(rB2in, block, 0, 0.45, 0, 0.45)\nsolid(sBall2, region, rB2in").replace(
        "group(gBall2, particles, region, rBall2, solid, sBall2)", "group(gBall2, particles, region, rBall2, solid, sBall2)"),
    "expression_functions": block((4, 4, 4), "musl").replace("0.5*a*(y-cy), 0.5*a*(z-cz))", "a*sqrt(y)*(x0-cx)/(1+z*z), (-a)*(z>cz)*y^2)"),
}


def _setup(lib, script, host):
    if host:
        os.environ["KML_HOST_POPULATE"] = "1"
    else:
        os.environ.pop("KML_HOST_POPULATE", None)
    try:
        e = Engine(lib)
        e.script(script)
        e.apply_initial_fixes()
        out = []
        for i in range(e.nsolids()):
            out.append({f: e.download(i, getattr(P, f)) for f in FIELDS})
        e.close()
    finally:
        os.environ.pop("KML_HOST_POPULATE", None)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(SCRIPTS))
def test_device_setup_is_bit_identical_to_the_host_path(cuda_lib, name):
    dev, host = _setup(cuda_lib, SCRIPTS[name], False), _setup(cuda_lib, SCRIPTS[name], True)
    assert len(dev) == len(host)
    for a, b in zip(dev, host):
        assert len(a["PTAG"]) == len(b["PTAG"]) > 0
        for f in FIELDS:
            assert a[f].tobytes() == b[f].tobytes(), (name, f, float(np.abs(a[f].astype(float) - b[f].astype(float)).max()))
    if "initial_velocity_particles" in SCRIPTS[name] and name != "expression_functions":
        assert any(np.abs(a["V"]).max() > 0 for a in dev), "the initial velocity fix did not act"


def test_expression_compiler_matches_the_interpreter(oracle_lib):
    """Host side of f1 (no GPU): the postfix program of an expression, evaluated in Python with IEEE doubles, equals the interpreter's
    per-particle result for +, -, *, / (the operations the device evaluates with explicitly rounded intrinsics)."""
    import ctypes as C
    lib = oracle_lib
    if not hasattr(lib, "kmlh_compile_expr"):
        pytest.skip("host library without kmlh_compile_expr")
    e = Engine(lib)
    e.script("a = 2.5e-4\ncx = 20\nb = 0.1\n")
    expr = "b-a*(x-cx)/(1.5+y0)*(-z)"
    n = C.c_int(0); ops = (C.c_int * 96)(); vals = (C.c_double * 96)()
    assert lib.kmlh_compile_expr(e.h, expr.encode(), C.byref(n), ops, vals) == 0
    rng = np.random.default_rng(0)
    for _ in range(20):
        pv = rng.uniform(0.5, 40, 6)
        st = []
        for i in range(n.value):
            op, v = ops[i], vals[i]
            if op == 0: st.append(np.float64(v))
            elif op == 1: st.append(np.float64(pv[int(v)]))
            elif op == 7: st[-1] = -st[-1]
            else:
                bb = st.pop(); aa = st.pop()
                st.append({2: aa + bb, 3: aa - bb, 4: aa * bb, 5: aa / bb}[op])
        for k, nm in enumerate(("x", "y", "z", "x0", "y0", "z0")):
            assert lib.kmlh_set_particle_var(e.h, nm.encode(), C.c_double(pv[k])) == 0
        assert float(st[0]) == e.var_eval(expr)
    e.close()
