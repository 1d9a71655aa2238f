"""Turn an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` log of bench.py into
per-kernel-group figures for exactly ONE training step (the launch sequence is periodic; the period is found from the
kernel names).  Usage: python profiles/summarize_ncu.py <ncu.csv> <out.json> [--group-prefix conv_tc_gather]"""
import collections
import csv
import json
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    d = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < 15:
            continue
        name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        d.setdefault(int(r[0]), {"name": name})[r[12]] = float(r[14].replace(",", ""))
    return list(d.values())


def period_of(names):
    for per in range(8, len(names) // 2 + 1):
        if names[:len(names) - per] == names[per:]:
            return per
    for per in range(8, len(names)):            # fewer than two full periods captured
        if names[:len(names) - per] == names[per:] and len(names) - per >= 8:
            return per
    return len(names)


def main():
    ks = load(sys.argv[1])
    names = [k["name"] for k in ks]
    marks = [i for i, n in enumerate(names) if "sgd_kernel" in n]       # the optimiser kernel closes a training step
    if len(marks) >= 3:
        step = ks[marks[1] + 1:marks[2] + 1]
        per = len(step)
    else:
        per = period_of(names)
        step = ks[:per]
    groups = collections.OrderedDict()
    for k in step:
        base = k["name"].split("<")[0]
        g = groups.setdefault(base, {"launches": 0, "time_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        g["launches"] += 1
        g["time_us"] += k.get("gpu__time_duration.sum", 0.0) / 1e3
        g["dram_read_bytes"] += k.get("dram__bytes_read.sum", 0.0)
        g["dram_write_bytes"] += k.get("dram__bytes_write.sum", 0.0)
    total = sum(g["time_us"] for g in groups.values())
    for g in groups.values():
        g["share_of_captured_time"] = g["time_us"] / total
        g["dram_bytes_per_launch"] = (g["dram_read_bytes"] + g["dram_write_bytes"]) / g["launches"]
        g["time_us_per_launch"] = g["time_us"] / g["launches"]
    out = {"source": sys.argv[1], "launches_in_capture": len(ks), "launches_per_step": per, "groups": groups,
           "note": "ncu serialises launches and runs them cold; shares, not absolute times, are comparable to bench.py"}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    for n, g in groups.items():
        print("%-28s x%-4d %8.1f us  %5.1f %%  %7.1f MB/launch DRAM" % (n, g["launches"], g["time_us"],
              100 * g["share_of_captured_time"], g["dram_bytes_per_launch"] / 1e6))


if __name__ == "__main__":
    main()
