"""Script-surface features beyond the shipped examples, checked against the UNMODIFIED reference binary: the computes
max_plastic_strain / average_velocity (src/compute_max_plastic_strain.cpp, src/compute_average_velocity.cpp), the gz dump
styles (src/dump_particle_gz.cpp, src/dump_grid_gz.cpp), and the output variables of fix velocity_particles
(src/fix_velocity_particles.cpp:228-300).  Build-container only (needs /root/reference and oracle/_ref); the back end is
the CPU oracle, the CUDA back end shares the same host code and runs the same cases in tests/test_parity_gpu.py."""
import glob
import gzip
import os
import shutil
import subprocess
import tempfile

import pytest

from cases import driven_tool, taylor_bar, tensile, two_disks
from conftest import ROOT
from test_shipped_examples import OUR_CLI, REF_BIN, log_rows

pytestmark = pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.isdir("/root/reference")), reason="needs /root/reference and oracle/_ref (build container only)")


def run_both(text):
    out = {}
    for name, exe in (("ref", REF_BIN), ("our", OUR_CLI)):
        d = tempfile.mkdtemp(prefix="kmlextra_")
        try:
            open(os.path.join(d, "in.mpm"), "w").write(text)
            p = subprocess.run([exe, "-i", "in.mpm"], cwd=d, capture_output=True, text=True, timeout=600)
            files = {os.path.basename(f): open(f, "rb").read() for f in sorted(glob.glob(os.path.join(d, "dump*")))}
            out[name] = (p.returncode, p.stdout + p.stderr, files)
        finally:
            shutil.rmtree(d, ignore_errors=True)
    return out["ref"], out["our"]


def same_log(ref, our):
    a, b = log_rows(ref), log_rows(our)
    assert len(a) >= 2 and len(a) == len(b), (len(a), len(b))
    for ra, rb in zip(a, b):
        assert len(ra) == len(rb), (ra, rb)
        for x, y in zip(ra, rb):
            assert abs(x - y) <= 2e-5 * max(abs(x), abs(y)) + 1e-12, (ra, rb)  # 6 printed digits; sums of ~1e-16 noise stay below 1e-12


@pytest.fixture(scope="module", autouse=True)
def _cli(oracle_lib):
    if not os.path.exists(OUR_CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)


def test_compute_average_velocity():
    text = taylor_bar("linear") + "compute(va, average_velocity, gBall1)\nlog_modify(custom, step, dt, time, va_x)\nlog(5)\nrun(20)\n"
    (rc_r, out_r, _), (rc_o, out_o, _) = run_both(text)
    assert rc_r == 0 and rc_o == 0, out_o[-400:]
    same_log(out_r, out_o)


def test_compute_max_plastic_strain_thermal():
    # the reference reads Solid::T unconditionally (src/compute_max_plastic_strain.cpp:79), so it only runs thermo-mechanical
    text = tensile(True) + "compute(Ep, max_plastic_strain, all)\nlog_modify(custom, step, dt, time, Ep_Epmax, Ep_Tmax)\nlog(20)\nrun(100)\n"
    (rc_r, out_r, _), (rc_o, out_o, _) = run_both(text)
    assert rc_r == 0 and rc_o == 0, out_o[-400:]
    same_log(out_r, out_o)


def test_gz_dumps_hold_the_reference_text():
    text = two_disks("musl") + ("dump(d1, all, particle/gz, 5, dump_p.*.LAMMPS.gz, x, y, z, vx, s11, seq, mass)\n"
                                "dump(d2, all, grid/gz, 5, dump_g.*.LAMMPS.gz, x, y, z, vx, vy, mass)\nrun(10)\n")
    (rc_r, _, files_r), (rc_o, out_o, files_o) = run_both(text)
    assert rc_r == 0 and rc_o == 0, out_o[-400:]
    assert len(files_r) == 4 and set(files_r) == set(files_o)
    for k, v in files_r.items():
        assert gzip.decompress(files_o[k]) == gzip.decompress(v), "gz dump %s differs from the reference's" % k


def test_fix_velocity_particles_publishes_the_reaction():
    text = driven_tool() + "log_modify(custom, step, dt, time, vtool_x, vtool_y)\nlog(10)\nrun(60)\n"
    (rc_r, out_r, _), (rc_o, out_o, _) = run_both(text)
    assert rc_r == 0 and rc_o == 0, out_o[-400:]
    same_log(out_r, out_o)


def test_translate_particles_external_force_delete_compute():
    # src/translate_particles.cpp, Group::external_force (src/group.cpp:410-462), Modify::delete_compute (src/modify.cpp:218-231)
    text = two_disks("musl") + ("region(rMove, block, 0, INF, 0, INF)\ntranslate_particles(sBall2, region, rMove, -0.05, 0.025*x0, 0)\n"
                                "compute(Ek, kinetic_energy, all)\nfe = external_force(gBall1, x)\nlog_modify(custom, step, dt, time, Ek, fe)\nlog(10)\n"
                                "dump(d1, all, particle, 20, dump_p.*.LAMMPS, x, y, z, x0, y0, vx)\nrun(40)\ndelete_compute(Ek)\n")
    (rc_r, out_r, files_r), (rc_o, out_o, files_o) = run_both(text)
    assert rc_r == 0 and rc_o == 0, out_o[-400:]
    same_log(out_r, out_o)
    assert len(files_r) == 2 and files_r == files_o


@pytest.mark.parametrize("tl", [False, True])
def test_delete_particles_keeps_the_reference_order(tl):
    # src/delete_particles.cpp:51-80: the dump lists the survivors in storage order, so byte-identical dumps pin the compaction order
    from cases import carved_disks
    text = carved_disks(tl) + ("compute(Ek, kinetic_energy, all)\nlog_modify(custom, step, dt, time, Ek)\nlog(10)\n"
                               "dump(d1, all, particle, 25, dump_p.*.LAMMPS, x, y, z, vx, vy, s11, mass)\nrun(50)\n")
    (rc_r, out_r, files_r), (rc_o, out_o, files_o) = run_both(text)
    assert rc_r == 0 and rc_o == 0, out_o[-400:]
    same_log(out_r, out_o)
    assert len(files_r) == 2 and files_r == files_o


def test_fix_check_solution():
    # src/fix_check_solution.cpp:104-200; the "all" group, where the reference initialises its total volume
    text = two_disks("musl") + ("fix(chk, check_solution, all, 0.1*time*(1+x0), 0.08*time)\n"
                                "log_modify(custom, step, dt, time, chk_s, chk_x, chk_y, chk_z)\nlog(5)\nrun(20)\n")
    (rc_r, out_r, _), (rc_o, out_o, _) = run_both(text)
    assert rc_r == 0 and rc_o == 0, out_o[-400:]
    same_log(out_r, out_o)


def test_composite_regions():
    # src/region_union.cpp, src/region_intersection.cpp, src/region_difference.cpp: a solid cut out of nested set operations
    text = """
E   = 1e+3
nu  = 0.3
rho = 1000
L   = 1
hL  = 0.5*L
method(ulmpm, FLIP, cubic-spline, 0.99)
N        = 20
cellsize = L/N
dimension(2,-hL, hL, -hL, hL, cellsize)
region(rA, cylinder, -0.15, -0.1, 0.2)
region(rB, block, -0.3, 0.1, -0.1, 0.2)
region(rC, cylinder, 0.2, 0.15, 0.15)
region(rHole, cylinder, -0.1, 0.0, 0.07)
region(rI, intersection, rA, rB)
region(rD, difference, rI, rHole)
region(rU, union, rD, rC)
material(mat1, linear, rho, E, nu)
solid(s1, region, rU, 2, mat1, cellsize, 0)
group(g1, particles, region, rU, solid, s1)
fix(v0, initial_velocity_particles, g1, 0.05, -0.1, NULL)
dt_factor(0.3)
compute(Ek, kinetic_energy, all)
log_modify(custom, step, dt, time, Ek)
log(10)
dump(d1, all, particle, 20, dump_p.*.LAMMPS, x, y, vx, s11)
run(20)
"""
    (rc_r, out_r, files_r), (rc_o, out_o, files_o) = run_both(text)
    assert rc_r == 0 and rc_o == 0, out_o[-400:]
    same_log(out_r, out_o)
    assert len(files_r) == 1 and files_r == files_o


@pytest.mark.parametrize("group,x0", [("all", 0.31), ("gBall2", 0.36)])
def test_fix_cuttingtool(group, x0):
    # src/fix_cutting_tool.cpp:118-285, both branches: "all" and a group restricted to one solid (where the reference adds the line's y
    # coefficient in place of its constant, :222-223 - mirrored)
    text = two_disks("musl", method="method(ulmpm, FLIP, cubic-spline, 0.99)") + (
        "xt = %g-0.1*time\nyt = %g-0.1*time\n"
        "fix(ftool, cuttingtool, %s, 0.05, xt, yt, 0, -0.1, -0.1, 0, xt+0.3, yt+0.05, xt+0.05, yt+0.3)\n"
        "log_modify(custom, step, dt, time, ftool_x, ftool_y)\nlog(10)\n"
        "dump(d1, all, particle, 30, dump_p.*.LAMMPS, x, y, vx, vy, s11)\nrun(30)\n" % (x0, x0, group))
    (rc_r, out_r, files_r), (rc_o, out_o, files_o) = run_both(text)
    assert rc_r == 0 and rc_o == 0, out_o[-400:]
    same_log(out_r, out_o)
    rows = log_rows(out_o)
    assert abs(rows[-1][3]) > 0.1, "the tool never touched the disk"
    assert len(files_r) == 1 and files_r == files_o
