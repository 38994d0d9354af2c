#!/usr/bin/env python
"""profiles/sass_digest.txt: per kernel of the shipped liblq.so, the SASS opcode counts that show what it is built from
(128-bit global loads/stores, TMA bulk/tensor copies, mbarrier waits, warp reductions, system-scope loads).  Runs
without a GPU (cuobjdump -sass).

    python tools/sass_digest.py > profiles/sass_digest.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "learned_quantization_b200", "csrc", "liblq.so")
WATCH = ["LDG.E.128", "LDG.E.EF.128", "LDG.E.STRONG.SYS", "STG.E.128", "STG.E.EF.128", "UTMALDG", "UTMASTG", "UBLKCP",
         "SYNCS", "REDUX", "SHFL", "MUFU.RCP", "MUFU.TANH", "BAR.SYNC", "ATOM", "RED.E", "MEMBAR", "CCTL", "ERRBAR",
         "LDS.128", "STS", "FMNMX", "HMMA", "UTCMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = re.findall(r"arch = (sm_\w+)", sass)
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(arch)} cubins, architectures {sorted(set(arch))}")
    print("# kernel | instructions | " + " ".join(WATCH))
    name, counts, total = None, collections.Counter(), 0
    rows = []

    def flush():
        if name is not None:
            rows.append((name, total, dict(counts)))
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            flush()
            name, counts, total = m.group(1), collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name is not None:
            total += 1
            op = m.group(1)
            for w in WATCH:
                if op.startswith(w):
                    counts[w] += 1
    flush()
    demangle = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    for (raw, tot, c), nice in sorted(zip(rows, demangle), key=lambda t: t[1]):
        nice = re.sub(r"\(.*", "", nice)[:110]
        hits = " ".join(f"{w}={c[w]}" for w in WATCH if c.get(w))
        print(f"{nice} | {tot} | {hits}")


if __name__ == "__main__":
    main()
