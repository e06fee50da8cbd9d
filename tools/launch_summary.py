"""Summarise an `ncu --metrics gpu__time_duration.sum[,smsp__cycles_active.avg] --csv` launch list: one step's kernels."""
import collections
import csv
import re
import sys


def main(path, per_step=None, skip=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = {}
    order = []
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        if i not in rows:
            rows[i] = dict(name=row["Kernel Name"].split("(")[0].replace("void ", "").replace("s3d::", ""),
                           grid=row["Grid Size"], block=row["Block Size"])
            order.append(i)
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            rows[i]["us"] = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        elif row["Metric Name"] == "smsp__cycles_active.avg":
            rows[i]["act"] = v
    seq = [rows[i] for i in order]
    # a step ends with k_sched_step
    # a step of the graph-replayed loop ends with the fused boundary kernel k_boundary<2, ...> (out head + scheduler + next in_conv)
    ends = [k for k, r in enumerate(seq) if re.match(r"k_boundary<(\(s3d::MODE\))?2[,>]", r["name"])]
    if len(ends) < 3:
        ends = [k for k, r in enumerate(seq) if r["name"].startswith("k_sched_step")]
    if len(ends) >= 3:
        step = seq[ends[-2] + 1: ends[-1] + 1]
    else:
        step = seq
    tot = collections.OrderedDict()
    for r in step:
        e = tot.setdefault(r["name"], [0, 0.0, 0.0])
        e[0] += 1
        e[1] += r.get("us", 0.0)
        e[2] += r.get("act", 0.0)
    print(f"{'kernel':22s} {'n':>3s} {'us':>8s} {'active kcyc':>12s}")
    for k, (n, us, act) in tot.items():
        print(f"{k:22s} {n:3d} {us:8.1f} {act / 1e3:12.1f}")
    print(f"{'TOTAL':22s} {len(step):3d} {sum(v[1] for v in tot.values()):8.1f} {sum(v[2] for v in tot.values()) / 1e3:12.1f}")
    if "-v" in sys.argv:
        try:
            for i, r in enumerate(step):
                print(i, r["name"], round(r.get("us", 0), 2), r["grid"], r["block"], round(r.get("act", 0)))
        except BrokenPipeError:
            pass


if __name__ == "__main__":
    main(sys.argv[1])
