"""Top stalled SASS lines + stall-reason totals from `ncu --page source --csv` output."""
import csv
import sys


def main(path, top=14):
    rows = list(csv.reader(open(path)))
    print(rows[0][1][:90] if len(rows[0]) > 1 else rows[0])
    hdr = rows[1]
    si, src = hdr.index("# Samples"), hdr.index("Source")
    data = [r for r in rows[2:] if len(r) > max(si, src)]
    num = lambda r: int(r[si]) if r[si].isdigit() else 0
    print("total samples", sum(num(r) for r in data))
    for r in sorted(data, key=lambda r: -num(r))[:top]:
        print(r[si].rjust(6), r[src][:110])
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {hdr[i]: 0 for i in cols}
    for r in data:
        for i in cols:
            if i < len(r) and r[i].isdigit():
                agg[hdr[i]] += int(r[i])
    print(sorted(agg.items(), key=lambda kv: -kv[1])[:6])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
