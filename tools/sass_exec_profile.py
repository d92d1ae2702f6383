"""Where a kernel spends its ISSUE slots: executed warp-instructions of one launch by opcode and by SASS address bucket.

    ncu -i REPORT.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > sass.csv
    python tools/sass_exec_profile.py sass.csv [bucket]
"""
import csv
import sys
from collections import Counter


def main(path, bucket=200):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[h]
    data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
    iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[iEx]) for r in data)
    print("instructions", len(data), "executed", tot)
    ops = Counter()
    for r in data:
        src = r[iSrc].strip()
        op = src.split()[1] if src.startswith("@") else src.split()[0]
        ops[op.split(".")[0]] += int(r[iEx])
    print("by opcode:", ", ".join("%s %.1f%%" % (o, 100.0 * c / tot) for o, c in ops.most_common(18)))
    for b in range(0, len(data), bucket):
        ex = sum(int(r[iEx]) for r in data[b:b + bucket])
        sm = sum(int(r[iS]) for r in data[b:b + bucket])
        if ex > tot * 0.01:
            print("  sass %5d..%5d  executed %5.1f%%  samples %6d   first: %s" % (b, b + bucket - 1, 100.0 * ex / tot, sm, data[b][iSrc].strip()[:60]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 200)
