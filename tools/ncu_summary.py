#!/usr/bin/env python
"""Key metrics of a .ncu-rep, or of its `ncu -i X.ncu-rep --page raw --csv` export (*_raw.csv):
one block per captured launch.  Runs here, no GPU needed."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "dram__sectors_read.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * m.get(unit, 1)


for path in sys.argv[1:]:
    if path.endswith(".csv"):
        out = "".join(l for l in open(path) if l.startswith('"'))
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
        name = d["Kernel Name"][0][:70]
        rd = to_bytes(*d["dram__bytes_read.sum"]); wr = to_bytes(*d["dram__bytes_write.sum"])
        t = float(d["gpu__time_duration.sum"][0].replace(",", ""))
        tu = d["gpu__time_duration.sum"][1]
        t_us = t * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(tu, 1)
        print(f"{path.split('/')[-1]}: {name}")
        print(f"   time {t_us:.1f} us  dram read {rd/1e9:.3f} GB  write {wr/1e9:.3f} GB  total {(rd+wr)/1e9:.3f} GB  "
              f"-> {(rd+wr)/t_us*1e-3:.0f} GB/s")
        for k in WANT[3:]:
            if k in d:
                print(f"   {k} = {d[k][0]} {d[k][1]}")
