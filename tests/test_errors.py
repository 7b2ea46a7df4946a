"""Error behaviour of the path.  The reference aborts inside its loops: a particle leaves the domain (src/solid.cpp:617-627),
det F <= 0 on an undamaged particle (src/solid.cpp:1208-1215).  Here the kernels raise bits of a device word that adjust_dt /
error_flags return, and the host stops the run with a message.  Checked twice: our host (oracle back end) stops at the same step as
the unmodified reference binary (build container only), and the CUDA engine stops at the same step as the oracle (-m gpu)."""
import os
import re
import subprocess

import pytest

from cases import block, two_disks
from conftest import ROOT
from test_shipped_examples import OUR_CLI, REF_BIN


def leaves_domain():
    return two_disks("musl", v=-3.0)


def collapses():  # volumetric compression at a rate that turns det(I + dt L) negative in the first full step
    s = block((4, 4, 4), "usl", fixed_dt=True, a=6.0)
    fix = [ln for ln in s.splitlines() if "initial_velocity_particles" in ln]
    assert len(fix) == 1
    return s.replace(fix[0], fix[0].replace("0.5*a*(y-cy)", "-a*(y-cy)").replace("0.5*a*(z-cz)", "-a*(z-cz)"))


CASES = {"leaves_domain": (leaves_domain, "left the domain", 1), "collapses": (collapses, "J<=0", 2)}


def last_step(out):
    rows = [int(ln.split()[0]) for ln in out.splitlines() if re.match(r"^\d+\s+[-+0-9.e]+\s+[-+0-9.e]+\s*$", ln)]
    return rows[-1] if rows else -1


@pytest.mark.parametrize("name", list(CASES))
def test_host_stops_where_the_reference_stops(oracle_lib, name, tmp_path):
    if not (os.path.exists(REF_BIN) and os.path.isdir("/root/reference")):
        pytest.skip("needs /root/reference and oracle/_ref (build container only)")
    if not os.path.exists(OUR_CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
    script, message, _ = CASES[name]
    (tmp_path / "in.mpm").write_text(script() + "log(1)\nrun(40)\n")
    ref = subprocess.run([REF_BIN, "-i", "in.mpm"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    our = subprocess.run([OUR_CLI, "-i", "in.mpm"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert ref.returncode != 0 and our.returncode != 0
    assert message in our.stdout + our.stderr
    assert last_step(ref.stdout) == last_step(our.stdout) >= 0, (ref.stdout[-300:], our.stdout[-300:])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_engine_raises_the_flag_at_the_same_step(cuda_lib, oracle_lib, name):
    from karamelo_b200.api import Engine, KmlError
    script, _, bit = CASES[name]
    steps = {}
    for label, lib in (("cuda", cuda_lib), ("oracle", oracle_lib)):
        e = Engine(lib)
        e.script(script())
        with pytest.raises(KmlError) as err:
            e.line("run(40)")
        steps[label] = e.state()["ntimestep"]
        if label == "cuda":
            assert "device error flags" in str(err.value) and int(re.search(r"flags (?:set: )?(\d+)", str(err.value)).group(1)) & bit, str(err.value)
        e.close()
    assert steps["cuda"] == steps["oracle"], steps
