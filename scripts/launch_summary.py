"""Per-kernel summary of one MD step from an ncu launch list (--metrics gpu__time_duration.sum --csv):
python scripts/launch_summary.py LAUNCHES.csv [OUT.txt]"""
import collections
import csv
import io
import re
import sys


def main():
    text = open(sys.argv[1]).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    names = [r["Kernel Name"] for r in rows]
    idx = [i for i, n in enumerate(names) if "baoab_pre" in n]
    a, b = idx[-2], idx[-1]
    agg = collections.OrderedDict()
    for r in rows[a:b]:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")[:64]
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("ns", "nsecond"):
            v /= 1000.0
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1000.0
        d = agg.setdefault(n, [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(v for _, v in agg.values())
    out = ["one MD step (between two baoab_pre launches) from an ncu launch list; per-launch times are cold-cache and "
           "serialised: compare SHARES with bench.py's kernels_ms_per_step, not absolutes",
           f"launches {b - a}   sum of kernel times {tot:.1f} us"]
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{n:66s} x{c:3d} {v:9.1f} us {100 * v / tot:5.1f}%")
    s = "\n".join(out) + "\n"
    print(s)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(s)


if __name__ == "__main__":
    main()
