#!/bin/bash
# One `ncu --set full` capture of each kernel of the split extraction path (148 clips x 300 frames).  usage: bash tools/gpu_ncu_split.sh <tag>
tag=${1:-q}
mkdir -p gpurun_out
for k in strip_sweep frame_regions region_variance; do
ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_${k}_$tag -f \
    python bench.py --steps 1 --warmup 1 --clips 148 --frames 300 --tracks 0 --no-motion --no-cpu-baseline > gpurun_out/ncu_full_${k}_$tag.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -4
