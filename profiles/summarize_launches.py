"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
    python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("odevit::<unnamed>::", "odevit::")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if "odevit::" in k)
    print(f"launches: {sum(v[0] for v in agg.values())}, total device time {tot / 1e3:.2f} ms, "
          f"libodevit kernels {ours / tot * 100:.1f} %\n")
    print("| share | launches | avg us | kernel |\n|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"| {v[1] / tot * 100:.2f} % | {v[0]} | {v[1] / v[0]:.1f} | `{k[:120]}` |")


if __name__ == "__main__":
    main(sys.argv[1])
