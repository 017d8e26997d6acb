mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --tracks 0 --no-motion --e2e-clips 8 --no-extras > gpurun_out/b.json 2> gpurun_out/b.err
python - <<PY
import json
d = json.loads(open("gpurun_out/b.json").read().strip().splitlines()[-1])
print("{:.2f} M frames/s, {:.2f} ms/step, kernels {}".format(d["value"] / 1e6, d["ms_per_step"], {k: round(v, 2) for k, v in d["roofline"]["kernel_times_ms"].items()}))
PY
