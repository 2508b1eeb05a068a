#!/usr/bin/env python3
"""Turns an `ncu --set full` report into the per-kernel summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/<name>.ncu-rep profiles/<round>_ncu_full_summary.csv [--traffic]

Reads the report with `ncu -i <rep> --page raw --csv` (ncu is on the CPU box; no GPU needed), keeps the metrics
the roofline discussion in DESIGN.md uses, one column per profiled launch.  With --traffic it also rewrites
profiles/idct_traffic.json (dram bytes per image of the IDCT/colour launch), which bench.py reports as
`roofline.traffic`."""
import csv
import io
import json
import os
import subprocess
import sys

METRICS = [
    "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
] + [f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio" for s in (
    "barrier", "long_scoreboard", "short_scoreboard", "math_pipe_throttle", "not_selected", "wait", "branch_resolving",
    "no_instruction", "mio_throttle", "lg_throttle")]

UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]].replace("jpgpu::", "").split("(")[0] for r in data]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        for m in METRICS:
            if m in col:
                w.writerow([m, units[col[m]]] + [r[col[m]] for r in data])
    print(f"{out}: {len(data)} launches: {names}")
    if "--traffic" in sys.argv:
        # profiles/idct_traffic.json: per workload key ("<W>x<H>_<subsampling>", given after --traffic) the DRAM bytes per
        # image of the IDCT/colour launch, stamped with the hash of the kernel sources it was measured on - bench.py only
        # reports the figure while that hash still matches (roofline.traffic is null for any other build).
        import hashlib
        key = sys.argv[sys.argv.index("--traffic") + 1]
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        h = hashlib.sha256()
        for f in ("jpgpu_kernels.cu", "jpgpu_core.h"):
            h.update(open(os.path.join(root, "jpeg_rust_b200", "csrc", f), "rb").read())
        path = os.path.join(root, "profiles", "idct_traffic.json")
        try:
            table = json.load(open(path))
            if not isinstance(table, dict) or "width" in table:
                table = {}
        except Exception:
            table = {}
        for r, name in zip(data, names):
            if "idct_colour_kernel" not in name:
                continue
            rd = float(r[col["dram__bytes_read.sum"]]) * UNIT_SCALE[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]]) * UNIT_SCALE[units[col["dram__bytes_write.sum"]]]
            grid = r[col["Grid Size"]].strip("()").split(",")
            images = int(grid[1])
            table[key] = {"kernel": name, "profile": os.path.basename(out), "images": images, "dram_bytes_read": rd,
                          "dram_bytes_write": wr, "dram_bytes_per_image": (rd + wr) / images,
                          "kernel_source_sha": h.hexdigest()[:16]}
            print(key, table[key])
        json.dump(table, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
