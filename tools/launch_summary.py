"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import re
import sys


def main(path, detail=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot, seq = {}, 0.0, []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name).replace("dcl::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        seq.append((name, v))
    print(f"total {tot:.1f} us over {len(seq)} launches")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v:10.1f} us {100 * v / tot:5.1f}% {n:4d}  {k}")
    if detail:
        for name, v in seq:
            if detail in name:
                print(f"   {v:9.1f}  {name}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
