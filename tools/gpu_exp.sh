#!/bin/bash
# A/B an experimental build (-DCPT_EXP=<n>) against the in-tree library: kernel times of one bench step each.  usage: bash tools/gpu_exp.sh <n> [<n> ...]
mkdir -p gpurun_out
for e in 0 "$@"; do
  lib=classifier-pipeline_b200/libcptrack.so
  if [ "$e" != 0 ]; then
    lib=/tmp/libcptrack_exp$e.so
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -DCPT_EXP=$e -Xcompiler -fPIC -shared -o $lib classifier-pipeline_b200/csrc/*.cu 2>/dev/null || { echo "build $e failed"; continue; }
  fi
  CPT_LIB=$(realpath $lib) python bench.py --no-cpu-baseline --tracks 0 --no-motion --e2e-clips 8 > gpurun_out/bench_exp$e.json 2> gpurun_out/bench_exp$e.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_exp$e.json").read().strip().splitlines()[-1])
print("EXP $e: {:.2f} M frames/s, {:.2f} ms/step, regions {}, kernels {}".format(d["value"] / 1e6, d["ms_per_step"], d["run_info"]["regions_found"], {k: round(v, 2) for k, v in d["roofline"]["kernel_times_ms"].items()}))
PY
done
