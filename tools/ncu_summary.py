#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, on the CPU box) into the few lines DESIGN.md / bench.py cite:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/ncu_x_r02.csv"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(r"^(gpu__time_duration\.sum|launch__(registers_per_thread|grid_size|block_size|occupancy_limit_\w+|waves_per_multiprocessor)|"
                  r"launch__shared_mem_per_block_\w+|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__inst_executed_pipe_(alu|fma|lsu|uniform|xu|adu|cbu)\.sum\.pct_of_peak_sustained_active|sm__inst_issued\.sum\.pct_of_peak_sustained_active|"
                  r"sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|sm__inst_executed\.avg\.per_cycle_active|smsp__inst_executed\.sum|"
                  r"sm__warps_active\.avg\.(pct_of_peak_sustained_active|per_cycle_active)|sm__cycles_active\.avg|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|smsp__sass_inst_executed_op_(shared_ld|shared_st|global_ld|global_st)\.sum|"
                  r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|lts__t_bytes\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed)$")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        w = csv.writer(f)
        for r in rows[2:]:
            name = dict(zip(hdr, r)).get("Kernel Name", "?")
            w.writerow(["kernel", name, ""])
            for h, u, v in zip(hdr, units, r):
                if KEEP.match(h):
                    try:
                        if float(v) == 0.0 and "stalled" in h:
                            continue
                    except ValueError:
                        pass
                    w.writerow([h, v, u])
    print(open(out).read())


if __name__ == "__main__":
    main()
