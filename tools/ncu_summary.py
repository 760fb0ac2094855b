"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family.

    python tools/ncu_summary.py gpurun_out/launches.csv [first_fraction_to_skip]

Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(b200::)?(\w+)(<[^>]*>)?", name)
    if name.startswith("at::") or "at::native" in name:
        m2 = re.search(r"(\w+Functor|\w+_kernel\w*|CatArrayBatchedCopy\w*|\w+Kernel\w*)", name)
        return "torch:" + (m2.group(1) if m2 else name[:60])
    if m:
        t = m.group(3) or ""
        return m.group(2) + t[:60]
    return name[:80]


def main():
    path = sys.argv[1]
    skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    rows = rows[int(len(rows) * skip):]
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in rows:
        k = short(r["Kernel Name"])
        tot[k] += float(r["Metric Value"]) / 1e6
        cnt[k] += 1
    total = sum(tot.values())
    print(f"{len(rows)} launches, {total:.2f} ms summed kernel time (cold-cache, serialised)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v:10.3f} ms {100 * v / total:6.2f}%  x{cnt[k]:<5d} avg {1e3 * v / cnt[k]:9.1f} us  {k}")


if __name__ == "__main__":
    main()
