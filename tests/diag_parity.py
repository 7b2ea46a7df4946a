#!/usr/bin/env python
"""Diagnostic (not a test): error growth of the CUDA engine against the oracle, step by step.

    python tests/diag_parity.py <case> [checkpoints...]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from cases import CASES  # noqa: E402
from common import FIELDS, rel  # noqa: E402
from karamelo_b200.api import Engine, load_host_library  # noqa: E402


def main():
    name = sys.argv[1]
    cps = [int(x) for x in sys.argv[2:]] or [1, 2, 3, 5, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100]
    script, is_tl, thermal, _ = CASES[name]
    cuda = load_host_library(None)
    oracle = load_host_library(os.path.join(os.path.dirname(HERE), "oracle", "_build", "libkml_host_oracle.so"))
    a, b = Engine(cuda), Engine(oracle)
    a.script(script)
    b.script(script)
    done = 0
    fields = FIELDS + (("T",) if thermal else ())
    print("%5s %12s " % ("step", "dt_rel") + " ".join("%10s" % f[:10] for f in fields[1:]))
    for cp in cps:
        a.line("run(%d)" % (cp - done))
        b.line("run(%d)" % (cp - done))
        done = cp
        sa, sb = a.snapshot(fields), b.snapshot(fields)
        row = []
        for f in fields[1:]:
            row.append(max(rel(x[f], y[f]) for x, y in zip(sa, sb)))
        da, db = a.state()["dt"], b.state()["dt"]
        print("%5d %12.3e " % (cp, abs(da - db) / abs(db)) + " ".join("%10.2e" % r for r in row), flush=True)


if __name__ == "__main__":
    main()
