mkdir -p gpurun_out
for lib in classifier-pipeline_b200/libcptrack.so build/libvar_64_1.so build/libvar_64_2.so build/libvar_128_2.so build/libvar_256_2.so build/libvar_64_4.so; do
  CPT_LIB=$(realpath $lib) timeout 300 python bench.py --no-cpu-baseline --tracks 0 --no-motion --e2e-clips 8 --no-extras > gpurun_out/b.json 2> gpurun_out/b.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/b.json").read().strip().splitlines()[-1])
print("$lib: {:.2f} ms/step, regions {}, kernels {}".format(d["ms_per_step"], d["run_info"]["regions_found"], {k[:12]: round(v, 2) for k, v in d["roofline"]["kernel_times_ms"].items()}))
PY
done
