"""Helpers for pinning the oracle against the real reference (test infrastructure).

* ``run_reference`` runs ``oracle/_ref/karamelo_ref`` (the unmodified reference compiled by
  ``oracle/Makefile``) on a script and returns the per-solid particle state it wrote
  through its own ``restart(N, file)`` command (binary doubles, reference
  ``src/solid.cpp:2841-2887``; matrices are Eigen column-major there).
* ``read_restart_solids`` locates the solid records inside a restart file.  The file starts
  with variable-length method / region / material blocks (``src/write_restart.cpp:51-88``);
  rather than re-implementing every writer, the solid header is found by its signature
  (np, np_local, nc, first and last particle tag), with np taken from the reference's own
  stdout (``np_local=`` lines printed by ``Solid::populate``, ``src/solid.cpp:2278``).
"""
import os
import re
import struct
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "karamelo_ref")


def _record_dtype(is_tl, temp):
    f = [("ptag", "<i8"), ("x0", "<f8", (3,)), ("x", "<f8", (3,)), ("v", "<f8", (3,)), ("sigma", "<f8", (3, 3)), ("strain_el", "<f8", (3, 3))]
    if is_tl:
        f.append(("vol0PK1", "<f8", (3, 3)))
    f += [("F", "<f8", (3, 3)), ("J", "<f8"), ("vol0", "<f8"), ("rho0", "<f8"), ("eps", "<f8"), ("epsdot", "<f8"), ("damage", "<f8"), ("damage_init", "<f8")]
    if temp:
        f.append(("T", "<f8"))
    f += [("ienergy", "<f8"), ("mask", "<i4")]
    return np.dtype(f)


def read_restart_solids(path, nps, is_tl, temp):
    buf = open(path, "rb").read()
    dt = _record_dtype(is_tl, temp)
    out, pos, first_tag = [], 0, 1
    for npart in nps:
        # np (bigint) is the solid's size at creation; np_local (int) follows and is smaller after delete_particles, whose swap-with-last
        # compaction also leaves the tags unordered (src/delete_particles.cpp:56-80) - so: every tag inside the solid's range, none twice
        sig = struct.pack("<q", npart)
        found = None
        o = buf.find(sig, pos)
        while o != -1:
            hdr = o - 96
            rec0 = hdr + 124
            nloc = struct.unpack_from("<i", buf, o + 8)[0] if o + 12 <= len(buf) else -1
            end = rec0 + nloc * dt.itemsize
            if hdr >= 0 and 0 < nloc <= npart and end <= len(buf):
                nc = struct.unpack_from("<i", buf, o + 12)[0]
                if nc in (0, 2, 4, 8):
                    tags = np.frombuffer(buf, dtype=dt, count=nloc, offset=rec0)["ptag"]
                    if tags.min() >= first_tag and tags.max() <= first_tag + npart - 1 and len(np.unique(tags)) == nloc and (nloc < npart or (tags[0] == first_tag and tags[-1] == first_tag + npart - 1)):
                        found = (hdr, rec0, end, nloc)
                        break
            o = buf.find(sig, o + 1)
        if found is None:
            raise RuntimeError("solid with np=%d (first tag %d) not found in %s" % (npart, first_tag, path))
        hdr, rec0, end, nloc = found
        rec = np.frombuffer(buf, dtype=dt, count=nloc, offset=rec0)
        d = {k: np.array(rec[k]) for k in dt.names}
        for k in ("sigma", "strain_el", "F", "vol0PK1"):
            if k in d:
                d[k] = np.ascontiguousarray(np.swapaxes(d[k], 1, 2))  # Eigen column-major -> row-major
        d["cellsize"] = struct.unpack_from("<d", buf, hdr + 116)[0]
        d["solidlo"] = np.array(struct.unpack_from("<3d", buf, hdr))
        d["solidhi"] = np.array(struct.unpack_from("<3d", buf, hdr + 24))
        out.append(d)
        pos, first_tag = end, first_tag + npart
    return out


_OUTPUT_CMDS = ("dump(", "log_modify(", "log(", "set_output(", "plot(", "save_plot(", "restart(", "run(", "run_time(",
                "run_until(", "run_while(")


def strip_output_commands(script):
    """Drop dump / log / plot / restart / run lines so a test can append its own."""
    keep = []
    for ln in script.splitlines():
        s = ln.split("#")[0].replace(" ", "")
        if not s.startswith(_OUTPUT_CMDS):
            keep.append(ln)
    return "\n".join(keep) + "\n"


def run_reference(script, steps, is_tl, temp=False, timeout=3600, keep_dir=None):
    """Run the reference on ``script`` + ``restart``/``run`` for ``steps`` steps; return (solids, stdout)."""
    if not os.path.exists(REF_BIN):
        raise FileNotFoundError(REF_BIN)
    d = keep_dir or tempfile.mkdtemp(prefix="kmlref_")
    text = script + "\nrestart(%d, ref-*.restart)\nrun(%d)\n" % (steps, steps)
    with open(os.path.join(d, "in.mpm"), "w") as f:
        f.write(text)
    p = subprocess.run([REF_BIN, "-i", "in.mpm"], cwd=d, capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError("reference failed:\n" + p.stdout[-3000:] + p.stderr[-2000:])
    nps = [int(m) for m in re.findall(r"^np_local=(\d+)", p.stdout, flags=re.M)]
    solids = read_restart_solids(os.path.join(d, "ref-%d.restart" % steps), nps, is_tl, temp)
    return solids, p.stdout


def rel_err(a, b):
    """max |a-b| / max(|b|, tiny) - the north_star's relative tolerance is on field magnitude."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale) if a.size else 0.0
