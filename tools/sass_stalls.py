"""Top stall locations of one kernel launch from an ncu report's SASS source page.

    ncu -i REPORT.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > sass.csv
    python tools/sass_stalls.py sass.csv [N]

Prints the share of each warp-stall reason over all samples and the N most-sampled instructions with their two
dominant reasons (the view that showed which loads / barriers a phase-serial kernel actually waits on).
"""
import csv
import sys


def main(path, n=40):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[h]
    data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
    iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[iS]) for r in data)
    print("samples", tot, "instructions", len(data), "executed", sum(int(r[iEx]) for r in data))
    for i in stall:
        s = sum(int(r[i]) for r in data)
        if s > tot * 0.01:
            print("  %-24s %5.1f%%" % (hdr[i], 100.0 * s / tot))
    top = sorted(enumerate(data), key=lambda t: -int(t[1][iS]))[:n]
    for idx, r in sorted(top):
        st = sorted(((int(r[i]), hdr[i][6:]) for i in stall if int(r[i]) > 0), reverse=True)[:2]
        print("%5d %6s %9s  %-72s %s" % (idx, r[iS], r[iEx], r[iSrc].strip()[:72], st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
