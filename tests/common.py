"""Shared comparison helpers for the parity tests."""
import os

import numpy as np

from karamelo_b200.api import Engine, P

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SYM = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
FIELDS = ("PTAG", "X", "V", "SIGMA", "FDEF", "EFF_PLASTIC_STRAIN", "EFF_PLASTIC_STRAIN_RATE", "DAMAGE", "DAMAGE_INIT")


def run_case(lib, script, steps, thermal=False, extra_fields=()):
    e = Engine(lib)
    e.script(script + "\nrun(%d)\n" % steps)
    snap = e.snapshot(FIELDS + (("T",) if thermal else ()) + tuple(extra_fields))
    st = e.state()
    e.close()
    return snap, st


def rel(a, b):
    """max |a - b| relative to the magnitude of the reference field (north_star tolerance)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    if a.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / scale


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = []
    for i in range(int(g["nsolids"])):
        d = {k: g["%s%d" % (k, i)] for k in ("ptag", "x", "v", "sigma", "F", "eps", "epsdot", "damage", "damage_init")}
        if "T%d" % i in g:
            d["T"] = g["T%d" % i]
        order = np.argsort(d["ptag"], kind="stable")  # Engine.snapshot sorts by tag; the reference's order is not ascending after delete_particles
        out.append({k: v[order] for k, v in d.items()})
    return out, g["log_last"]


def compare_to_golden(snap, golden, tol, floors=None):
    """Returns {field: worst relative error}; asserts tags are identical and errors within tol."""
    worst = {}
    assert len(snap) == len(golden)
    for s, r in zip(snap, golden):
        assert s["PTAG"].shape == r["ptag"].shape and (s["PTAG"] == r["ptag"]).all(), "particle counts / tags differ"
        pairs = {
            "x": (s["X"], r["x"]), "v": (s["V"], r["v"]),
            "sigma": (np.stack([s["SIGMA"][:, a, b] for a, b in SYM], 1), r["sigma"]),
            "F": (s["FDEF"].reshape(len(r["ptag"]), 9), r["F"]),
            "eps": (s["EFF_PLASTIC_STRAIN"], r["eps"]), "epsdot": (s["EFF_PLASTIC_STRAIN_RATE"], r["epsdot"]),
            "damage": (s["DAMAGE"], r["damage"]), "damage_init": (s["DAMAGE_INIT"], r["damage_init"]),
        }
        if "T" in r and "T" in s:
            pairs["T"] = (s["T"], r["T"])
        for k, (a, b) in pairs.items():
            worst[k] = max(worst.get(k, 0.0), rel(a, b))
    bad = _beyond(worst, tol, "sigma", "F")
    assert not bad, "fields beyond tolerance %g: %s (all: %s)" % (tol, bad, worst)
    return worst


FLIP_SIGMA_TOL = 1e-9


def _beyond(worst, tol, sigma_key, f_key):
    """Fields beyond the tolerance.  One allowance, for comparisons of the CUDA engine at 1e-10: the reference keeps a node in a particle's
    neighbour list only `if (wf != 0)` (src/ulmpm.cpp:252-263) and whether the outermost spline weight rounds to zero depends on the last bit of
    the particle position, which differs between two summation orders (the CUDA node sums are atomic).  Such a flip drops |v| dw dt (at most
    ~1e-11) of F, which a stiff material shows as up to (bulk modulus / stress level) times that in the stress - the reference algorithm
    disagrees with ITSELF by that much when only its summation order changes (tests/test_weight_zero_skip.py).  So when
    every other field, F included, is inside the tolerance, the stress alone is held to 1e-9 instead of 1e-10."""
    bad = {k: v for k, v in worst.items() if v > tol}
    if tol >= 1e-10 and set(bad) == {sigma_key} and bad[sigma_key] <= FLIP_SIGMA_TOL and worst.get(f_key, 0.0) <= tol:
        print("note: stress at %.2e with every other field inside %g - a neighbour-membership flip (see tests/common.py _beyond)" % (bad[sigma_key], tol))
        return {}
    return bad


def compare_snaps(a, b, tol):
    worst = {}
    assert len(a) == len(b)
    for s, r in zip(a, b):
        assert (s["PTAG"] == r["PTAG"]).all(), "particle counts / tags differ"
        for k in s:
            if k == "PTAG":
                continue
            worst[k] = max(worst.get(k, 0.0), rel(s[k], r[k]))
    bad = _beyond(worst, tol, "SIGMA", "FDEF")
    assert not bad, "fields beyond tolerance %g: %s (all: %s)" % (tol, bad, worst)
    return worst


def rel_either(got, ref_a, ref_b):
    """max over elements of min(|got - a|, |got - b|), relative to the magnitude of a.  For cases with marginal neighbour-membership
    events (test_weight_zero_skip.py): a and b are the oracle with and without the reference's `wf != 0` test, and an element may
    follow either branch of a coin flip the reference itself takes on the last bit of a particle position."""
    g, a, b = np.asarray(got, float), np.asarray(ref_a, float), np.asarray(ref_b, float)
    if g.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(a))), 1e-300)
    return float(np.max(np.minimum(np.abs(g - a), np.abs(g - b)))) / scale


def oracle_both_memberships(lib, script, steps, fields, all_solids=False):
    """The oracle run twice: reference semantics (nodes whose weight rounds to 0 are dropped, src/ulmpm.cpp:252-263) and with those nodes kept."""
    out = []
    for keep in (False, True):
        if keep:
            os.environ["KML_ORACLE_KEEP_ZERO_WEIGHT"] = "1"
        try:
            e = Engine(lib)
            e.script(script + "\nrun(%d)\n" % steps)
            snaps = e.snapshot(fields)
            out.append((snaps if all_solids else snaps[0], e.state()))
            e.close()
        finally:
            os.environ.pop("KML_ORACLE_KEEP_ZERO_WEIGHT", None)
    return out


def permute_particles(e, seed=1, solid=0):
    """Shuffle the particle arrays of a freshly populated solid in place (tags keep their identity): nothing in the step may depend on
    the order particles were created in (the reference rebuilds its neighbour lists every step, src/ulmpm.cpp:140-156)."""
    fields = [P.PTAG, P.X, P.X0, P.V, P.MASS, P.VOL0, P.VOL, P.RHO0, P.MASK]
    data = [e.download(solid, f) for f in fields]
    perm = np.random.default_rng(seed).permutation(len(data[0]))
    for f, a in zip(fields, data):
        e.upload(solid, f, np.ascontiguousarray(a[perm]))
    return perm
