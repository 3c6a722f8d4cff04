"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def summarize(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.OrderedDict()
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            if d["Metric Unit"] in ("nsecond", "ns"):
                v /= 1000.0
            agg.setdefault(d["Kernel Name"], []).append(v)
    return agg


if __name__ == "__main__":
    for k, v in summarize(sys.argv[1]).items():
        print("%-64s n=%3d mean %9.1f us  min %9.1f" % (k[:64], len(v), sum(v) / len(v), min(v)))
