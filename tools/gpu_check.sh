#!/bin/bash
# One GPU-box pass: parity tests, the bench line, the ncu launch list and one full capture of the kernel.
# usage (from the repo root, under gpurun): bash tools/gpu_check.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; cat gpurun_out/bench_ref_$tag.json
python tools/phase_timing.py 148 300 0 > gpurun_out/phases_$tag.txt 2>&1; cat gpurun_out/phases_$tag.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --clips 148 --frames 200 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:extract_clips -c 1 -o gpurun_out/prof_$tag -f \
    python bench.py --steps 1 --warmup 1 --clips 148 --frames 100 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out | tail -12
