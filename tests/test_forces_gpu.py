"""Particle forces f_p = a_p m_p through the CUDA engine (Group::internal_force / external_force, src/group.cpp:340-470;
the reference stores a_p during grid_to_points, src/solid.cpp:576-635).  The engine derives a_p in registers and keeps it only
after kml_keep_particle_acceleration, which the host driver calls when a script defines such a variable."""
import numpy as np
import pytest

from cases import tensile, two_disks
from karamelo_b200.api import Engine, KmlError, P

pytestmark = pytest.mark.gpu

CASES = {
    "ul_two_disks": two_disks("musl") + "fe = external_force(gBall1, x)\nfi = internal_force(gBall2, y)\n",
    "tl_tensile": tensile(False) + "group(gp, particles, region, region2, solid, solid1)\nfe = external_force(gp, x)\nfi = internal_force(gp, x)\n",
    "ul_block_cell_kernels": None,  # filled below: the cell G2P kernel is replaced by the generic one while accelerations are kept
}


def _script(name):
    if CASES[name] is not None:
        return CASES[name]
    from cases import block
    return block((6, 6, 6), "musl") + "fe = external_force(gall, x)\nfi = internal_force(gall, z)\n"


@pytest.mark.parametrize("name", list(CASES))
def test_forces_match_oracle(cuda_lib, oracle_lib, name):
    out = []
    for lib in (cuda_lib, oracle_lib):
        e = Engine(lib)
        e.script(_script(name) + "run(30)\n")
        out.append((e.var("fe"), e.var("fi"), e.download(0, P.A), e.download(0, P.V_UPDATE), e.download(0, P.F), e.download(0, P.PTAG)))
        e.close()
    (fe, fi, a, vu, f, tag), (fe_o, fi_o, a_o, vu_o, f_o, tag_o) = out
    assert (tag == tag_o).all()
    scale = max(np.abs(a_o).max(), 1e-300)
    assert np.abs(a - a_o).max() <= 1e-9 * scale, np.abs(a - a_o).max() / scale   # a = sum w (v~ - v) / dt amplifies node rounding by 1 / dt
    assert np.abs(vu - vu_o).max() <= 1e-10 * max(np.abs(vu_o).max(), 1e-300)
    assert np.abs(f - f_o).max() <= 1e-9 * max(np.abs(f_o).max(), 1e-300)
    fs = max(np.abs(f_o).sum(), 1e-300)
    assert abs(fe - fe_o) <= 1e-9 * fs and abs(fi - fi_o) <= 1e-9 * fs, (fe, fe_o, fi, fi_o)


def test_forces_must_be_requested_before_the_first_step(cuda_lib):
    e = Engine(cuda_lib)
    e.script(two_disks("musl") + "run(2)\n")
    with pytest.raises(KmlError):
        e.line("fe = external_force(gBall1, x)")
    e.close()
