"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: last forward only."""
import collections
import csv
import re
import sys


def main(path, per_forward):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(r["Kernel Name"], float(r["Metric Value"])) for r in rows]
    last = names[-per_forward:]
    tot = sum(v for _, v in last)
    print(f"{len(rows)} launches captured; last forward = {len(last)} launches, {tot / 1e6:.3f} ms serialized (cold cache)")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in last:
        k = re.sub(r"\(.*", "", n).replace("void dirb200::<unnamed>::", "").replace("dirb200::<unnamed>::", "")
        agg[k][0] += 1
        agg[k][1] += v
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v / 1e6:9.3f} ms {100 * v / tot:5.1f}%  x{c:3d}  {k[:90]}")
    print("-- launches > 0.1 ms, in order")
    for i, (n, v) in enumerate(last):
        if v / 1e6 > 0.1:
            print(f"{i:4d} {v / 1e6:.3f} ms", re.sub(r"\(.*", "", n)[-60:])


if __name__ == "__main__":
    try:
        main(sys.argv[1], int(sys.argv[2]))
    except BrokenPipeError:  # piped into head
        pass
