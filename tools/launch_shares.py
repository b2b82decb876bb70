"""Kernel shares of an ncu launch list (--metrics gpu__time_duration.sum --csv): markdown table on stdout.

    python tools/launch_shares.py gpurun_out/r2_launches_bench_step.csv
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("<unnamed>::", "").replace("void ", "")
    return name.replace("(bool)", "")


def main(path):
    tot, cnt = defaultdict(float), defaultdict(int)
    with open(path) as fh:
        rows = [r for r in fh if r.startswith('"')]
    for r in csv.DictReader(rows):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        k = short(r["Kernel Name"])
        tot[k] += ns
        cnt[k] += 1
    total = sum(tot.values())
    print("| kernel | launches | total ms | share | us / launch |\n|---|---|---|---|---|")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"| `{k}` | {cnt[k]} | {tot[k] / 1e6:.3f} | {100 * tot[k] / total:.1f} % | {tot[k] / cnt[k] / 1e3:.1f} |")
    print(f"\nTotal {total / 1e6:.1f} ms over {sum(cnt.values())} launches.")


if __name__ == "__main__":
    main(sys.argv[1])
