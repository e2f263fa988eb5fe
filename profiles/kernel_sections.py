"""One line per distinct kernel from an `ncu --page raw --csv` export of a multi-kernel capture.
Usage: ncu -i capture.ncu-rep --page raw --csv > raw.csv; python profiles/kernel_sections.py raw.csv"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    seen = collections.OrderedDict()
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        m = re.search(r"(?:obia::)?(\w+)(<[^>]*>)?\(", name)
        short = (m.group(1) + (m.group(2) or "")) if m else name[:40]
        seen.setdefault(short, []).append(r)

    def f(r, k, default=0.0):
        try:
            return float(r[col[k]].replace(",", ""))
        except (KeyError, ValueError):
            return default

    print(f"{'kernel':44s} {'xN':>4s} {'ms':>9s} {'dram%':>6s} {'TB/s':>6s} {'issue%':>7s} {'regs':>5s} {'warps%':>7s} "
          f"{'GB read':>8s} {'GB written':>10s}")
    for short, rs in seen.items():
        r = rs[len(rs) // 2]
        t = f(r, "gpu__time_duration.sum")
        unit = rows[1][col["gpu__time_duration.sum"]]
        ms = t / 1e6 if unit in ("ns", "nsecond") else t / 1e3 if unit in ("us", "usecond") else t
        scale = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9}
        if "dram__bytes_read.sum" in col:
            rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
            rd *= scale.get(rows[1][col["dram__bytes_read.sum"]], 1e-9)
            wr *= scale.get(rows[1][col["dram__bytes_write.sum"]], 1e-9)
        else:
            # section-only captures carry the rate, not the byte counts: total traffic goes in the "read" column
            u = rows[1][col["dram__bytes.sum.per_second"]]
            rate = f(r, "dram__bytes.sum.per_second") * {"Tbyte/s": 1e3, "Gbyte/s": 1.0, "Mbyte/s": 1e-3,
                                                         "Tbyte/second": 1e3, "Gbyte/second": 1.0,
                                                         "Mbyte/second": 1e-3}.get(u, 1e-9)
            rd, wr = rate * ms / 1e3, 0.0
        print(f"{short[:44]:44s} {len(rs):4d} {ms:9.4f} {f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{(rd + wr) / (ms / 1e3) / 1e3 if ms else 0:6.2f} "
              f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f} {int(f(r, 'launch__registers_per_thread')):5d} "
              f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):7.1f} {rd:8.3f} {wr:10.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
