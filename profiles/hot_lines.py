"""Aggregate `ncu --page source --csv --print-source cuda` output: top source lines by stall samples."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    out = []
    fname = None
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Name":
            fname = r[1].split("/")[-1]
            hdr = None
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            samples = int(d.get("# Samples", "0") or 0)
            inst = int(d.get("Instructions Executed", "0") or 0)
        except ValueError:
            continue
        if samples or inst:
            out.append((samples, inst, fname, d["Line No"], d["Source"].strip()[:110]))
    tot = sum(o[0] for o in out) or 1
    toti = sum(o[1] for o in out) or 1
    out.sort(reverse=True)
    print(f"# {path}: total samples {tot}, total warp-instructions {toti}")
    for s, i, f, ln, src in out[:top]:
        print(f"{100.0 * s / tot:5.1f}% smp {100.0 * i / toti:5.1f}% inst  {f}:{ln:>5s}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
