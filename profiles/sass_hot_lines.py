"""Join an `ncu --page source --csv` (SASS view, per-instruction samples) with the cubin's line table
(`nvdisasm --print-line-info`) and print the hottest CUDA source lines.

usage: sass_hot_lines.py <ncu_sass_csv> <cubin> <kernel-substring> [top]
"""
import collections
import csv
import re
import subprocess
import sys


def line_table(cubin, kern):
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    table, cur, infn = {}, None, False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            infn = kern in m.group(1)
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m and cur:
            table[int(m.group(1), 16)] = cur
    return table


def main(path, cubin, kern, top=40):
    table = line_table(cubin, kern)
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hdr]
    ia, isamp, iinst = H.index("Address"), H.index("# Samples"), H.index("Instructions Executed")
    base = None
    agg = collections.defaultdict(lambda: [0, 0])
    for r in rows[hdr + 1:]:
        if len(r) <= iinst:
            continue
        if r[ia] == "Address":      # a second launch of the same kernel starts a new section
            base = None
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        key = table.get(a - base, ("?", 0))
        agg[key][0] += int(r[isamp] or 0)
        agg[key][1] += int(r[iinst] or 0)
    tot = sum(v[0] for v in agg.values()) or 1
    toti = sum(v[1] for v in agg.values()) or 1
    srcs = {}
    print(f"# {kern}: {tot} stall samples, {toti} warp-instructions; top source lines")
    for (f, ln), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            try:
                srcs[f] = open(f"fast-livo_b200/csrc/{f}").read().splitlines()
            except OSError:
                srcs[f] = []
        src = srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""
        print(f"{100.0 * s / tot:5.1f}% smp {100.0 * i / toti:5.1f}% inst  {f}:{ln:<5d} {src}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 40)
