#!/usr/bin/env python
"""Join two outputs of tools/kmat_rate.py (labels `old` and `new`, same box, libraries swapped between the runs) into one
markdown table:  python tools/kmat_ab_table.py old.md new.md"""
import sys


def rows(path):
    d = {}
    for line in open(path):
        if line.startswith("| old") or line.startswith("| new"):
            c = [x.strip() for x in line.strip().strip("|").split("|")]
            d[(c[1], c[2], c[3], c[4])] = (float(c[5]), float(c[6]), float(c[7]), c[8])
    return d


def main():
    old, new = rows(sys.argv[1]), rows(sys.argv[2])
    print("| kernel | Real | Ns x Nt | row pitch | old ms | old GB/s | new ms | new GB/s | new / old | new, frac of HBM copy peak | same bits |\n|---|---|---|---|---|---|---|---|---|---|---|")
    for k, a in old.items():
        b = new[k]
        print("| %s | %s | %s | %s | %.4f | %.0f | %.4f | %.0f | %.2f | %.3f | %s |" % (k + (a[0], a[1], b[0], b[1], b[1] / a[1], b[2], "yes" if a[3] == b[3] else "NO")))


if __name__ == "__main__":
    main()
