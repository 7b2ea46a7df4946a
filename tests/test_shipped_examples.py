"""Drop-in check of the script surface: EVERY example file the reference ships runs UNMODIFIED - only the run length is
shortened - through our host driver.  Where the unmodified reference runs the file, the log table we print must equal
the reference's; where the reference itself stops with an error (half of the shipped files predate the current
parser), we must stop too.  Build-container only: it reads /root/reference and runs oracle/_ref (both absent on the
GPU box, where these tests skip); the back end here is the CPU oracle, the CUDA back end shares the same host code."""
import glob
import os
import re
import shutil
import subprocess
import tempfile

import pytest

from conftest import ROOT

REF_EXAMPLES = "/root/reference/examples"
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "karamelo_ref")
OUR_CLI = os.path.join(ROOT, "oracle", "_build", "kml_oracle")
# the BASELINE.json configs as shipped get a longer run
LONG = {"two-disks.mpm": 40, "Taylor-bar/cylindrical/ULMPM/cylindrical_taylor_bar.mpm": 12,
        "Tensile_with_damage/Bernstein/inputfile": 300, "Bouncing_balls/TLMPM/FLIP/bouncing_balls2.mpm": 40}


def shipped():
    files = sorted(glob.glob(os.path.join(REF_EXAMPLES, "**", "*.mpm"), recursive=True) + glob.glob(os.path.join(REF_EXAMPLES, "**", "inputfile"), recursive=True))
    return [os.path.relpath(f, REF_EXAMPLES) for f in files]


def log_rows(text):
    rows = []
    for ln in text.splitlines():
        parts = ln.split()
        if len(parts) >= 3 and re.fullmatch(r"\d+", parts[0]):
            try:
                rows.append([float(x) for x in parts])
            except ValueError:
                pass
    return rows


def run(exe, src, text, keep_dumps=None):
    d = tempfile.mkdtemp(prefix="kmlship_")
    try:
        for aux in os.listdir(os.path.dirname(src)):  # meshes etc. next to the script
            p = os.path.join(os.path.dirname(src), aux)
            if os.path.isfile(p) and os.path.getsize(p) < (8 << 20) and aux != os.path.basename(src):
                os.symlink(p, os.path.join(d, aux))
        open(os.path.join(d, "in.mpm"), "w").write(text)
        p = subprocess.run([exe, "-i", "in.mpm"], cwd=d, capture_output=True, text=True, timeout=600)
        if keep_dumps is not None:
            for f in sorted(glob.glob(os.path.join(d, "dump*"))):
                keep_dumps[os.path.basename(f)] = open(f, "rb").read()
        return p.returncode, p.stdout + p.stderr
    finally:
        shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("rel", shipped() or ["(no reference tree)"])
def test_shipped_script_behaves_like_the_reference(oracle_lib, rel):
    src = os.path.join(REF_EXAMPLES, rel)
    if not (os.path.exists(src) and os.path.exists(REF_BIN)):
        pytest.skip("needs /root/reference and oracle/_ref (build container only)")
    if not os.path.exists(OUR_CLI):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, capture_output=True)
    n = LONG.get(rel, 6)
    text = re.sub(r"(?m)^run(_time|_until|_while)?\(.*$", "run(%d)" % n, open(src, errors="replace").read())
    if rel in LONG:
        text = re.sub(r"(?m)^(log|set_output)\(.*$", r"\1(%d)" % max(n // 4, 1), text)
    dumps_ref, dumps_our = {}, {}
    rc_ref, out_ref = run(REF_BIN, src, text, dumps_ref)
    rc_our, out_our = run(OUR_CLI, src, text, dumps_our)
    if rc_ref != 0:
        assert rc_our != 0, "the reference rejects this file, we accept it:\n" + out_ref[-400:]
        return
    assert rc_our == 0, "the reference runs this file, we stop:\n" + out_our[-600:]
    ref, ours = log_rows(out_ref), log_rows(out_our)
    assert len(ref) >= 1 and len(ref) == len(ours), (len(ref), len(ours))
    for a, b in zip(ref, ours):
        assert len(a) == len(b), (a, b)
        for x, y in zip(a, b):
            assert abs(x - y) <= 2e-5 * max(abs(x), abs(y)) + 1e-30, (a, b)  # the log prints 6 significant digits
    # the dump files (LAMMPS text, what users' OVITO scripts read) are the same files, byte for byte
    lammps = {k: v for k, v in dumps_ref.items() if k.endswith(".LAMMPS")}
    assert set(lammps) == {k for k in dumps_our if k.endswith(".LAMMPS")}, (sorted(lammps), sorted(dumps_our))
    for k, v in lammps.items():
        assert dumps_our[k] == v, "dump file %s differs from the reference's" % k
