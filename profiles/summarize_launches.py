"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv`) into per-kernel totals and shares.
usage: python profiles/summarize_launches.py launches.csv [first_id last_id] > summary.csv"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i + 1
            break
    ki, vi, ii, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID"), h.index("Metric Unit")
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    tot = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi or not (lo <= int(r[ii]) <= hi):
            continue
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui] in ("ns", "nsecond") else v / 1e3 if r[ui] in ("us", "usecond") else v
        name = re.sub(r"\(.*$", "", r[ki]).replace("void ", "").strip()
        c, t = tot.get(name, (0, 0.0))
        tot[name] = (c + 1, t + ms)
    total = sum(t for _, t in tot.values())
    print("kernel,launches,total_ms,share")
    for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%s,%d,%.4f,%.3f" % (name, c, t, t / total))
    print("TOTAL,%d,%.4f,1.000" % (sum(c for c, _ in tot.values()), total))


if __name__ == "__main__":
    main()
