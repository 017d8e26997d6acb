#!/bin/bash
# Per-kernel durations of one full-size bench step (ncu launch list: cold-cache, serialised).  usage: bash tools/gpu_launches.sh <tag>
tag=${1:-q}
mkdir -p gpurun_out
KERNELS='regex:extract_clips|strip_sweep|frame_scalars|frame_regions|frame_components|region_variance'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -c 40 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --tracks 0 --no-motion --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/launches_$tag.csv") if l.startswith('"')))
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
for k, v in d.items():
    print("{:28s} n={:3d} mean {:12.3f} {}".format(k, len(v), sum(v) / len(v), rows[1][ui]))
PY
