#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED reference.

Runs every case of tests/cases.py through oracle/_ref/karamelo_ref (the reference sources
compiled by oracle/Makefile with the single-rank MPI + fixed-size Eigen shims), reads the
particle state the reference wrote with its own restart() command and stores it as
<case>.npz.  Needs /root/reference (build container only); the fixtures travel.

    python tests/golden/make_golden.py [case ...]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from cases import CASES  # noqa: E402
from refio import run_reference  # noqa: E402

SYM = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]


def main(names):
    for name in names:
        script, is_tl, thermal, steps = CASES[name]
        solids, out = run_reference(script, steps, is_tl, thermal)
        data = {"nsolids": len(solids), "steps": steps}
        for i, s in enumerate(solids):
            data["ptag%d" % i] = s["ptag"]
            data["x%d" % i] = s["x"]
            data["v%d" % i] = s["v"]
            data["sigma%d" % i] = np.stack([s["sigma"][:, a, b] for a, b in SYM], axis=1)
            data["F%d" % i] = s["F"].reshape(len(s["ptag"]), 9)
            for k in ("eps", "epsdot", "damage", "damage_init"):
                data["%s%d" % (k, i)] = s[k]
            if thermal:
                data["T%d" % i] = s["T"]
        # the last log line of the reference carries dt and time at 6 digits; keep for a sanity check
        last = [ln for ln in out.splitlines() if ln[:1].isdigit()][-1].split()
        data["log_last"] = np.array([float(x) for x in last[:3]])
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        nps = [len(s["ptag"]) for s in solids]
        print("%-26s np=%s steps=%d max|eps|=%.3g max damage=%.3g -> %s (%d kB)" % (
            name, nps, steps, max(s["eps"].max() for s in solids), max(s["damage"].max() for s in solids), os.path.basename(path), os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
