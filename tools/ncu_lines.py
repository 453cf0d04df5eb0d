"""Per-source-line view of one kernel: `ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:X
--launch-count 1 > src.csv`, then `python tools/ncu_lines.py src.csv [file-substring] [top]`: warp instructions and stall
samples per CUDA source line (inlined code is attributed to the line it was written on)."""
import csv
import sys


def main(path, want="", top=60):
    rows = list(csv.reader(open(path)))
    cur, hdr, agg = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr and len(r) == len(hdr) and r[0].isdigit():
            key = (cur, int(r[0]))
            a = agg.setdefault(key, [0, 0, r[1]])
            a[0] += int(r[iI])
            a[1] += int(r[iS])
    ti = sum(a[0] for a in agg.values())
    ts = sum(a[1] for a in agg.values())
    print("warp-inst", ti, "samples", ts)
    items = [(k, a) for k, a in agg.items() if want in k[0]]
    for k, a in sorted(items, key=lambda x: -x[1][1])[:top]:
        print(f"{k[0].split('/')[-1]:>18}:{k[1]:5d} inst {100 * a[0] / ti:5.1f}% samp {100 * a[1] / ts:5.1f}%  {a[2].strip()[:100]}")
    return agg, ti, ts


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 60)
