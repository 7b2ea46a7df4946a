"""Restart files: `restart(N, file-*.restart)` / `write_restart(file)` of our host write the reference's binary layout
(src/write_restart.cpp:51-88 and the write_restart member of every class it calls: Update, Domain, regions, Material with its EOS /
strength / damage / temperature lists, Solid, Group, Modify and the fixes).  With the CPU oracle behind the host the state is
bit-identical to the reference's, so the whole FILE must be: every parity case is run for 20 steps through the unmodified reference
binary and through our CLI and the two restart files are compared byte for byte (the version string of the header is set to the
reference build's).  Build-container only (needs /root/reference and oracle/_ref)."""
import os
import subprocess

import pytest

from cases import CASES
from conftest import ROOT
from test_shipped_examples import OUR_CLI, REF_BIN

pytestmark = pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.isdir("/root/reference")), reason="needs /root/reference and oracle/_ref (build container only)")
ENV = dict(os.environ, KML_RESTART_VERSION="oracle-ref-build")  # Version::GIT_SHA1 of oracle/ref_version.cpp


def run(exe, d, text):
    os.makedirs(d)
    with open(os.path.join(d, "in.mpm"), "w") as f:
        f.write(text)
    p = subprocess.run([exe, "-i", "in.mpm"], cwd=d, capture_output=True, text=True, timeout=600, env=ENV)
    assert p.returncode == 0, (exe, p.stdout[-300:], p.stderr[-300:])


@pytest.fixture(scope="module", autouse=True)
def _cli(oracle_lib):
    if not os.path.exists(OUR_CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)


@pytest.mark.parametrize("name", list(CASES))
def test_restart_file_is_the_reference_file(name, tmp_path):
    text = CASES[name][0] + "\nrestart(10, rst-*.restart)\nrun(20)\n"
    run(REF_BIN, str(tmp_path / "ref"), text)
    run(OUR_CLI, str(tmp_path / "our"), text)
    for step in (10, 20):
        a = open(tmp_path / "ref" / ("rst-%d.restart" % step), "rb").read()
        b = open(tmp_path / "our" / ("rst-%d.restart" % step), "rb").read()
        assert len(a) == len(b) and a == b, "restart file of step %d differs from the reference's (first difference at byte %d of %d)" % (
            step, next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b))), len(a))


def test_write_restart_command_only_records_the_name(tmp_path):
    # WriteRestart::command stores the file name and returns (src/write_restart.cpp:36-48): neither program writes a file
    text = CASES["c1_two_disks_musl"][0] + "\nrun(5)\nwrite_restart(snap-*.restart)\n"
    run(REF_BIN, str(tmp_path / "ref"), text)
    run(OUR_CLI, str(tmp_path / "our"), text)
    assert not [f for d in ("ref", "our") for f in os.listdir(tmp_path / d) if f.endswith(".restart")]


CONTINUE = ("read_restart(rst-20.restart)\nrestart(20, cont-*.restart)\n"
            "dump(d1, all, particle, 20, dump_p.*.LAMMPS, x, y, z, vx, vy, s11, s12, ep, damage)\nrun(20)\n")
# cases whose restart files the reference itself can continue from: every updated-Lagrangian case without a rigid material and without CPDI
# particle domains (the layout holds neither; total-Lagrangian runs stop in TLMPM because Domain::np_local is not restored); the reference's
# fix temperature_nodes cannot be re-created from a restart file either (it looks its group up by the name "restart",
# src/fix_temperature_nodes.cpp:38)
READABLE = [n for n, c in CASES.items() if not c[1] and "rigid" not in c[0] and "cpdi" not in c[0] and "velocity_particles" not in n
            and "temperature_nodes" not in c[0]]


@pytest.mark.parametrize("name", READABLE)
def test_read_restart_continues_like_the_reference(name, tmp_path):
    """The reference writes a restart file at step 20; both programs read THAT file, run 20 more steps and write a dump and a new restart
    file: both must be byte-identical (src/read_restart.cpp:36-83 and the read_restart member of every class it calls)."""
    run(REF_BIN, str(tmp_path / "ref"), CASES[name][0] + "\nrestart(20, rst-*.restart)\nrun(20)\n")
    os.makedirs(tmp_path / "our")
    with open(tmp_path / "our" / "rst-20.restart", "wb") as f:
        f.write(open(tmp_path / "ref" / "rst-20.restart", "rb").read())
    for exe, d in ((REF_BIN, "ref"), (OUR_CLI, "our")):
        with open(tmp_path / d / "cont.mpm", "w") as f:
            f.write(CONTINUE)
        p = subprocess.run([exe, "-i", "cont.mpm"], cwd=tmp_path / d, capture_output=True, text=True, timeout=600, env=ENV)
        assert p.returncode == 0, (exe, p.stdout[-300:], p.stderr[-300:])
    for fn in ("cont-40.restart", "dump_p.40.LAMMPS"):
        assert open(tmp_path / "ref" / fn, "rb").read() == open(tmp_path / "our" / fn, "rb").read(), fn


@pytest.mark.parametrize("name", ["c3_tensile_bernstein", "c4_balls_minpen", "x_rigid_ul_linear_musl"])
def test_read_restart_stops_where_the_reference_stops(name, tmp_path):
    """Total-Lagrangian runs (Domain::np_local is not restored, src/tlmpm.cpp:91-99) and rigid materials (the writer stores nothing, the
    reader expects a density, src/material.cpp:478-484,598-601) cannot be continued by the reference; neither by us."""
    run(REF_BIN, str(tmp_path / "ref"), CASES[name][0] + "\nrestart(20, rst-*.restart)\nrun(20)\n")
    for exe in (REF_BIN, OUR_CLI):
        with open(tmp_path / "ref" / "cont.mpm", "w") as f:
            f.write(CONTINUE)
        p = subprocess.run([exe, "-i", "cont.mpm"], cwd=tmp_path / "ref", capture_output=True, text=True, timeout=600, env=ENV)
        assert p.returncode != 0, exe
