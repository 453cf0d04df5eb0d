"""SASS evidence per kernel of libsphx.so: instruction count and the mnemonics that show packed fp32 (FFMA2 / FADD2 /
FMUL2), asynchronous copies (LDGSTS = cp.async, UBLKCP = cp.async.bulk, UTMALDG = TMA tensor copies), shared-memory and
global loads, MUFU. usage: python tools/sass_counts.py [sphexa_b200/libsphx.so] > profiles/rNN_sass_counts.txt"""
import collections
import re
import subprocess
import sys

KEYS = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP",
        "UTMALDG", "SYNCS", "BAR", "SHFL", "DFMA", "DADD", "DMUL", "F2F", "ATOM", "RED", "LDL", "STL"]


def main(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    name, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            counts[name], total[name] = collections.Counter(), 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            total[name] += 1
            op = m.group(1)
            if op in KEYS:
                counts[name][op] += 1
    print(f"# {lib}: static SASS instruction counts per kernel (cuobjdump -sass)")
    for k, c in counts.items():
        print(f"{k}: total {total[k]}  " + "  ".join(f"{op} {c[op]}" for op in KEYS if c[op]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "sphexa_b200/libsphx.so")
