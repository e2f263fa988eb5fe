"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py` per kernel.
Usage: python profiles/launch_summary.py launches.csv STEPS_IN_THE_RUN > summary.txt"""
import collections
import csv
import re
import sys


def main(path, steps):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        ms = {"ns": v / 1e6, "us": v / 1e3, "usecond": v / 1e3, "ms": v, "msecond": v}.get(r[iu], v * 1e3)
        m = re.search(r"obia::(\w+)", r[ik])
        name = m.group(1) if m else "ATen/other: " + r[ik].split("(")[0][-60:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    own = sum(v[1] for k, v in agg.items() if not k.startswith("ATen"))
    print(f"{len(rows) - 1} launches over {steps} steps; own kernels: {own / steps:.3f} ms of kernel time per step")
    print("(per-launch times under ncu are cold-cache and serialised: compare SHARES with the live CUDA-event numbers)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / steps:9.3f} ms/step  x {v[0] / steps:6.1f}  {100 * v[1] / own:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
