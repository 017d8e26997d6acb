#!/bin/bash
# Quick GPU pass while iterating on a kernel: parity tests, phase cycles, the bench line.  usage: bash tools/gpu_quick.sh <tag>
tag=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$tag.log
python tools/phase_timing.py 148 300 0 > gpurun_out/phases_$tag.txt 2>&1; cat gpurun_out/phases_$tag.txt
python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
