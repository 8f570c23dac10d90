#!/usr/bin/env python
"""Print the roofline-relevant metrics of every kernel in an .ncu-rep (read here on the CPU box with `ncu -i`).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [...]   (or the `ncu -i ... --page raw --csv` export made on the GPU box)"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]


def main():
    for path in sys.argv[1:]:
        out = open(path).read() if path.endswith(".csv") else subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"# {path}: no kernels"); continue
        hdr, units = rows[0], rows[1]
        print(f"# {path}")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            print(f"## {name[:150]}")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print(f"   {w:75s} {r[i]:>16s} {units[i]}")
            tens = [h for h in hdr if "tensor" in h and h not in WANT]
            for h in tens[:12]:
                i = hdr.index(h)
                print(f"   {h:75s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
