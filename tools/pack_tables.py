#!/usr/bin/env python3
"""Pack the reference's Butcher-tableau data files into ONE text file.

The reference loads `<label><q>.txt` (format: m, n, then m*n numbers, one per line;
reference main.cc:599-656).  north_star says the stage basis change uses "T/T^-1 from
tables/", and SURVEY 2.4(6) says the numbers must be consumed verbatim (regenerating the
Radau-IIA eigen-decompositions would change eigenvector scaling / ordering).  The data
(numbers only, no code) is therefore re-packed losslessly into

    dealii_spirk_b200/tables/butcher_tables.txt

one record per line:  `<label> <q> <m> <n> v0 v1 ...`  (values printed with repr(), i.e.
round-trip exact).  Only the labels the hot path reads are packed (main.cc:676-681,
1778-1786) plus A and L for the table-identity tests.

Run in the build container only (needs /root/reference):
    python tools/pack_tables.py
"""
import os
import re
import sys

SRC = "/root/reference/tables"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "dealii_spirk_b200", "tables", "butcher_tables.txt")
LABELS = ["A", "A_inv", "L", "T", "T_inv", "b_vec_", "c_vec_", "D_vec_",
          "T_re", "T_im", "T_inv_re", "T_inv_im", "D_vec_re_", "D_vec_im_"]


def main():
    recs = []
    for fn in sorted(os.listdir(SRC)):
        m = re.fullmatch(r"(.*?)(\d+)\.txt", fn)
        if not m or m.group(1) not in LABELS:
            continue
        label, q = m.group(1), int(m.group(2))
        toks = open(os.path.join(SRC, fn)).read().split()
        rows, cols = int(toks[0]), int(toks[1])
        vals = [float(t) for t in toks[2:]]
        assert len(vals) == rows * cols, fn
        recs.append((label, q, rows, cols, vals))
    recs.sort(key=lambda r: (LABELS.index(r[0]), r[1]))
    with open(DST, "w") as f:
        f.write("# packed from the reference's tables/*.txt by tools/pack_tables.py; "
                "format: label q m n values...\n")
        for label, q, rows, cols, vals in recs:
            f.write(" ".join([label, str(q), str(rows), str(cols)] + [repr(v) for v in vals]) + "\n")
    print("wrote", DST, len(recs), "records")


if __name__ == "__main__":
    sys.exit(main())
