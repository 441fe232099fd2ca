"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (us, share)."""
import collections
import csv
import sys


def main(path, steps):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0][-60:]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if v != v:                      # "nan": a launch the profiler could not time
            continue
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"total {tot:.1f} us over {steps} steps -> {tot / steps:.1f} us/step (cold-cache, serialised)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:62s} n={a[0]:4d} per-step={a[1] / steps:9.1f} us avg={a[1] / a[0]:8.1f} us share={a[1] / tot:.3f}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
