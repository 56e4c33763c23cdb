"""Summarises an `ncu --page source --csv --print-source cuda,sass` export by source line and by named line ranges.
usage: python tools/ncu_regions.py export.csv [file.cu:first-last=name ...]"""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    out, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1]
            continue
        if r and r[0] == "Function Name":
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and r and r[0] != "":
            try:
                out.append((cur.split("/")[-1], int(r[0]), int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]),
                            int(r[hdr.index("Thread Instructions Executed")]), r[1][:110]))
            except ValueError:
                continue
    return out


if __name__ == "__main__":
    data = load(sys.argv[1])
    ts, ti = sum(o[2] for o in data), sum(o[3] for o in data)
    print(f"total: {ts} samples, {ti} warp instructions")
    for spec in sys.argv[2:]:
        loc, name = spec.split("=")
        f, rng = loc.split(":")
        a, b = (int(v) for v in rng.split("-"))
        sel = [o for o in data if o[0] == f and a <= o[1] <= b]
        s, i, t = sum(o[2] for o in sel), sum(o[3] for o in sel), sum(o[4] for o in sel)
        print(f"{name:24s} samples {100 * s / ts:5.1f}%  inst {100 * i / ti:5.1f}%  active lanes {t / max(i, 1):4.1f}")
    if len(sys.argv) == 2:
        for o in sorted(data, key=lambda o: -o[3])[:40]:
            print(f"{100 * o[2] / ts:5.1f}% smp {100 * o[3] / ti:5.1f}% inst lanes {o[4] / max(o[3], 1):4.1f} {o[0]}:{o[1]}  {o[5]}")
