"""Summarise an ncu report (the `--set full` capture of one kernel) into the few numbers DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [frames_in_launch [traffic.json]] > profiles/X_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def main(path, frames=None, json_out=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("kernel:", name)
        got = {}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
                print("  {:90s} {:>18s} {}".format(h, v, u))
                got[h] = (v, u)
        try:
            def num(k):
                v, u = got[k]
                x = float(v.replace(",", ""))
                return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
            traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
            dur = num("gpu__time_duration.sum")
            print("  derived: dram traffic {:.4g} B, duration {:.4g} s, {:.1f} GB/s".format(traffic, dur, traffic / dur / 1e9))
            if frames:
                print("  derived: {:.0f} B of DRAM traffic per frame ({} frames in the launch); algorithmic 134400 B/frame".format(traffic / frames, frames))
                if json_out:
                    import json

                    json.dump({"source": path, "kernel": name, "frames_in_captured_launch": frames, "dram_bytes": traffic,
                               "dram_bytes_per_frame": traffic / frames, "duration_s": dur}, open(json_out, "w"), indent=1)
        except Exception as e:  # pragma: no cover
            print("  (derived figures unavailable: {})".format(e))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None, sys.argv[3] if len(sys.argv) > 3 else None)
