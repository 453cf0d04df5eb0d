"""Group the SASS of one kernel (ncu --page source --csv) into regions of similar execution count: share of
instructions and stall samples per region. usage: python tools/ncu_regions.py src.csv [dump_from dump_to]"""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
    iA = hdr.index("Address")
    first = data[0][iA]
    ends = [k for k, r in enumerate(data) if r[iA] == first]
    if len(ends) > 1:
        data = data[:ends[1]]
    return hdr, data


def main(path, dump=None):
    hdr, data = load(path)
    iS, iSm, iI = (hdr.index(k) for k in ("Source", "# Samples", "Instructions Executed"))
    ti = sum(int(r[iI]) for r in data)
    ts = sum(int(r[iSm]) for r in data)
    print("sass", len(data), "warp-inst", ti, "samples", ts)
    c = 0
    while c < len(data):
        e, base = c, int(data[c][iI])
        while e < len(data) and abs(int(data[e][iI]) - base) <= 0.35 * max(base, 1):
            e += 1
        i = sum(int(r[iI]) for r in data[c:e])
        s = sum(int(r[iSm]) for r in data[c:e])
        if i > 0.004 * ti or s > 0.004 * ts:
            print(f"{c:5d}-{e:5d} n={e - c:4d} exec/inst={base:10d} inst {100 * i / ti:5.1f}% samples {100 * s / ts:5.1f}%")
        c = e
    if dump:
        for k in range(*dump):
            r = data[k]
            print(f"{k:6d} {int(r[iSm]):7d} {int(r[iI]):10d}  {r[iS].strip()[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else None)
