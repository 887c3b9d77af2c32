"""Build profiles/r02_traffic.json (the `roofline.traffic` of the bench line) from `ncu -i X.ncu-rep --page raw --csv`
exports: dram__bytes_read.sum + dram__bytes_write.sum per launch of each persistent kernel, per workload.

  python profiles/make_traffic.py C2=gpurun_out/r2_ncu_c2_raw.csv C3=gpurun_out/r2_ncu_c3_raw.csv ...
"""
import csv
import json
import os
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main(args):
    out = {"source": "ncu --set full --clock-control none, profiles/r02_ncu_*_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum per launch)"}
    for a in args:
        wl, path = a.split("=")
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        col = {n: i for i, n in enumerate(hdr)}
        per = {}
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("<")[0].split("(")[0].replace("void ", "").replace("flb::", "").strip()
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[col[m]].replace(",", "")) * UNIT.get(units[col[m]], 1)
            per.setdefault(name, []).append(tot)
        out[wl] = {k: int(sum(v) / len(v)) for k, v in per.items()}
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_traffic.json")
    json.dump(out, open(p, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
