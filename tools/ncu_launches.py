#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (the CSV written by
`ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --clock-control none
 --csv --log-file X.csv <command>`): launches, average time and share of the total per kernel, plus
DRAM bytes when they were collected.  Runs anywhere (no GPU, no ncu needed).

    python tools/ncu_launches.py profiles/r1_laplace2d_launches.csv [--markdown]
"""
import argparse
import collections
import csv

TIME = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def load(path, by_grid=False):
    """-> ordered {kernel name: {"n": launches, "us": total time, "rd": bytes, "wr": bytes}};
    by_grid keeps launches of one kernel with different grid sizes apart (a full-matrix launch
    and the row pieces of the pipelined host-vector path are the same kernel)"""
    per_launch = collections.OrderedDict()
    with open(path, newline="") as f:
        for r in csv.reader(f):
            if len(r) < 15 or not r[0].isdigit():
                continue
            key = (int(r[0]), r[4] + (f"  grid {r[8]}" if by_grid else ""))
            per_launch.setdefault(key, {})[r[12]] = (float(r[14].replace(",", "")), r[13])
    agg = collections.OrderedDict()
    for (_, name), m in per_launch.items():
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        if "gpu__time_duration.sum" in m:
            v, u = m["gpu__time_duration.sum"]
            a["us"] += v * TIME.get(u, 1.0)
        for metric, field in (("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr")):
            if metric in m:
                v, u = m[metric]
                a[field] += v * BYTES.get(u, 1.0)
    return agg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--markdown", action="store_true")
    ap.add_argument("--by-grid", action="store_true", help="keep different grid sizes of one kernel apart")
    args = ap.parse_args()
    agg = load(args.csv, args.by_grid)
    total = sum(a["us"] for a in agg.values()) or 1.0
    with_bytes = any(a["rd"] or a["wr"] for a in agg.values())
    if args.markdown:
        print("| launches | avg us | share |" + (" DRAM read GB | DRAM write GB |" if with_bytes else "") + " kernel |")
        print("|---|---|---|" + ("---|---|" if with_bytes else "") + "---|")
    for name, a in agg.items():
        cols = [str(a["n"]), f"{a['us'] / a['n']:.1f}", f"{100 * a['us'] / total:.1f} %"]
        if with_bytes:
            cols += [f"{a['rd'] / a['n'] / 1e9:.3f}", f"{a['wr'] / a['n'] / 1e9:.3f}"]
        if args.markdown:
            print("| " + " | ".join(cols) + f" | `{name[:110] if not args.by_grid else name[:70] + name[name.rfind('  grid'):]}` |")
        else:
            print("  ".join(c.rjust(10) for c in cols) + "  " + name[:110])


if __name__ == "__main__":
    main()
