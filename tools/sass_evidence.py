#!/usr/bin/env python
"""Greps the SASS of the in-tree libmuopdb_gpu.so for the instructions DESIGN.md claims, kernel by kernel.
usage: python tools/sass_evidence.py > profiles/sass_evidence_<round>.txt      (needs cuobjdump; no GPU)
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops,
REDUX = warp reduce, MATCH = match.any, ATOMS = shared-memory atomics."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "muopdb_b200", "libmuopdb_gpu.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "LDS", "STS", "LDG", "PRMT", "IADD3", "REDUX", "MATCH", "ATOMS",
             "SHFL", "BAR.SYNC", "FFMA", "FADD", "FMUL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    sha = subprocess.run(["sha256sum", LIB], capture_output=True, text=True).stdout.split()[0]
    print(f"# cuobjdump -sass muopdb_b200/libmuopdb_gpu.so   (sha256 {sha[:16]}...)")
    print("# kernel | total SASS instructions | " + " ".join(MNEMONICS))
    cur, counts, total = None, None, 0
    rows = []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                rows.append((cur, total, counts))
            cur, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            total += 1
            op = m.group(1)
            for k in MNEMONICS:
                if op == k or op.startswith(k + ".") or (k == "BAR.SYNC" and op.startswith("BAR.SYNC")):
                    counts[k] += 1
    if cur:
        rows.append((cur, total, counts))
    dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    for (name, total, c), d in sorted(zip(rows, dem), key=lambda x: x[1]):
        d = re.sub(r"\(.*", "", d).replace("void ", "")
        print(f"{d} | {total} | " + " ".join(f"{k}={c[k]}" for k in MNEMONICS if c[k]))


if __name__ == "__main__":
    sys.exit(main())
