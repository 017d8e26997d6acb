"""Aggregate an ncu source page (--print-source cuda,sass --csv) per CUDA source line."""
import csv
import sys
from collections import defaultdict


def main(path, top=50):
    rows = list(csv.reader(open(path)))
    fname = None
    hdr = None
    agg = defaultdict(lambda: [0.0, 0.0, ""])
    stall_cols = []
    stalls = defaultdict(lambda: defaultdict(float))
    cur = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            stall_cols = [h for h in r if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < 8:
            continue
        if r[0] != "":
            cur = (fname, r[0])
            agg[cur][2] = r[1].strip()[:100]
        if cur is None:
            continue
        try:
            ins = float(r[hdr["Instructions Executed"]] or 0)
            smp = float(r[hdr["# Samples"]] or 0)
        except (ValueError, KeyError):
            continue
        agg[cur][0] += ins
        agg[cur][1] += smp
        for sc in stall_cols:
            try:
                stalls[cur][sc] += float(r[hdr[sc]] or 0)
            except ValueError:
                pass
    ti = sum(v[0] for v in agg.values())
    ts = sum(v[1] for v in agg.values())
    print("total warp-instructions {:.4g}  samples {:.4g}".format(ti, ts))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        st = sorted(stalls[k].items(), key=lambda kv: -kv[1])[:3]
        st = " ".join("{}={:.0f}%".format(a.replace("stall_", ""), 100 * b / max(v[1], 1)) for a, b in st)
        print("{:5.1f}% smp {:5.1f}% inst  {}:{:>4}  {:<80} | {}".format(100 * v[1] / ts, 100 * v[0] / ti, k[0][:14], k[1], v[2][:80], st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 50)
