"""Summarise `ncu -i rep --page source --csv --kernel-name regex:X` output: hot SASS regions by samples/instructions."""
import csv
import sys


def main(path, chunk=32, thresh=0.015, dump=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != "Address"]
    iA, iS, iSm, iI = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed"))
    first = data[0][iA]
    ends = [k for k, r in enumerate(data) if r[iA] == first]
    if len(ends) > 1:
        data = data[:ends[1]]
    tot_s = sum(int(r[iSm]) for r in data)
    tot_i = sum(int(r[iI]) for r in data)
    print("total samples", tot_s, "total inst", tot_i, "n sass", len(data))

    def op(t):
        t = t.split()
        return (t[1] if t[0].startswith('@') else t[0]).split('.')[0]

    for c in range(0, len(data), chunk):
        seg = data[c:c + chunk]
        s = sum(int(r[iSm]) for r in seg)
        i = sum(int(r[iI]) for r in seg)
        if s > tot_s * thresh or i > tot_i * thresh:
            ops = " ".join(sorted(set(op(r[iS]) for r in seg)))[:130]
            print(f"{c:6d} samples {100 * s / tot_s:5.1f}% inst {100 * i / tot_i:5.1f}%  {ops}")
    if dump:
        a, b = dump
        for k in range(a, b):
            r = data[k]
            print(f"{k:6d} {int(r[iSm]):7d} {int(r[iI]):10d}  {r[iS].strip()[:100]}")


if __name__ == "__main__":
    d = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else None
    main(sys.argv[1], dump=d)
