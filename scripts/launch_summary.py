"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: python scripts/launch_summary.py file.csv"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt, mx = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("ofl::<unnamed>::", "").replace("ofl::", "")[:60]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        ms = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v
        tot[name] += ms
        cnt[name] += 1
        mx[name] = max(mx[name], ms)
    total = sum(tot.values())
    print("| kernel | launches | total ms | share % | max ms |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / total:.1f} | {mx[k]:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1])
