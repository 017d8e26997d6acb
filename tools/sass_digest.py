"""Opcode histogram per kernel of libcptrack.so (cuobjdump -sass): the evidence for what the sm_100a build really contains
-- UBLKCP (1-D TMA bulk copy), SYNCS (mbarrier), IDP.2A (dp2a), REDUX / CREDUX (warp reductions), VIADDMNMX, ... -- and for
what it does not (no UTC*MMA / UTMALDG: nothing here is a contraction or a tiled tensor copy).
usage: python tools/sass_digest.py > profiles/r2_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "classifier-pipeline_b200", "libcptrack.so")
NOTABLE = ["UBLKCP", "UBLKPF", "UTMALDG", "SYNCS", "IDP", "REDUX", "CREDUX", "VIADDMNMX", "VIMNMX3", "VIMNMX", "VIADD", "ATOMS", "LDS", "STS", "LDG",
           "STG", "MUFU", "DFMA", "DADD", "DMUL", "IMAD", "FADD", "SHFL", "VOTE", "BAR", "UTCHMMA", "UTCQMMA", "HMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    print("cuobjdump -sass {}: architectures {}".format(os.path.relpath(LIB, ROOT), arch))
    print("kernel: total instructions | notable opcodes (static counts)")
    for name, c in kernels.items():
        notable = " ".join("{}={}".format(k, c[k]) for k in NOTABLE if c[k])
        print("{:34s} {:6d} | {}".format(name.replace("cpt::", ""), sum(c.values()), notable))
    total = collections.Counter()
    for c in kernels.values():
        total.update(c)
    print("\nwhole library: " + " ".join("{}={}".format(k, total[k]) for k in NOTABLE if total[k]))
    print("tensor-core / tensor-map opcodes (UTC*MMA, HMMA, UTMALDG): {}".format(sum(v for k, v in total.items() if k.startswith("UTC") or k in ("HMMA", "UTMALDG"))))


if __name__ == "__main__":
    main()
