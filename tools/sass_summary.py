#!/usr/bin/env python
"""cuobjdump -sass of the built libkml.so, summarised per headline kernel: instruction count, opcode histogram, and the mnemonics that prove
what the kernel uses (B200_PROFILING.md: UBLKCP / UTMALDG = TMA, SYNCS = mbarrier, LDGSTS = cp.async, REDG = fp64 reductions to the grid).

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "karamelo_b200", "lib", "libkml.so")
WANT = ["k_p2g_cell3", "k_g2p_cell", "k_stress_cell", "k_g2p_cell_tma", "k_g2p_cell_bulk", "k_cell_count", "k_cell_fill", "k_permute", "k_grid_update", "k_lattice_fill", "k_set_particles_expr"]
KEY = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "REDG", "ATOMG", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "BAR", "MUFU", "HMMA", "UTC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, kernels = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and cur:
            kernels[cur].append(m.group(1).strip())
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("SASS summary of %s (sm_100a), %d kernels in the library\n" % (os.path.relpath(LIB, ROOT), len(kernels)))
    for (name, ins), nice in zip(kernels.items(), demangle):
        if not any(w in name for w in WANT):
            continue
        ops = collections.Counter()
        for i in ins:
            t = i.split()
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            ops[op.split(".")[0]] += 1
        keys = {k: sum(v for o, v in ops.items() if o.startswith(k)) for k in KEY}
        print(nice[:150])
        print("  instructions %d | %s" % (len(ins), "  ".join("%s %d" % (k, v) for k, v in keys.items() if v)))
        print("  top opcodes: %s\n" % ", ".join("%s %d" % kv for kv in ops.most_common(12)))


if __name__ == "__main__":
    main()
