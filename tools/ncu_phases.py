"""Phase-level share of instructions / stall samples from an ncu cuda,sass source CSV.
usage: ncu_phases.py cs.csv name:lo-hi name:lo-hi ...  (line ranges of extract_kernel.cu)"""
import csv
import sys
from collections import defaultdict


def main(path, specs):
    rows = list(csv.reader(open(path)))
    fname = hdr = cur = None
    per, smp = defaultdict(float), defaultdict(float)
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or len(r) < 8 or r[0] == "":
            continue
        cur = (fname, int(r[0]))
        try:
            per[cur] += float(r[hdr["Instructions Executed"]] or 0)
            smp[cur] += float(r[hdr["# Samples"]] or 0)
        except ValueError:
            pass
    ti, ts = sum(per.values()), sum(smp.values())
    print("warp-instructions {:.4g}, samples {:.4g}".format(ti, ts))
    for spec in specs:
        name, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        i = sum(v for (f, l), v in per.items() if f.startswith("extract_kernel") and lo <= l <= hi)
        s_ = sum(v for (f, l), v in smp.items() if f.startswith("extract_kernel") and lo <= l <= hi)
        print("{:24s} inst {:5.1f}%  time {:5.1f}%".format(name, 100 * i / ti, 100 * s_ / ts))
    i = sum(v for (f, l), v in per.items() if not f.startswith("extract_kernel"))
    s_ = sum(v for (f, l), v in smp.items() if not f.startswith("extract_kernel"))
    print("{:24s} inst {:5.1f}%  time {:5.1f}%".format("inlined helpers (.cuh)", 100 * i / ti, 100 * s_ / ts))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
