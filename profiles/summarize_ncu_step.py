"""Per-kernel table of ONE training step from `ncu --csv --metrics ...` (profiles/r02_collect.sh, step 3): for every kernel
group (function name + grid x block, i.e. one conv shape) the launch count, summed time, share of the step, and the
time-weighted tensor-pipe %, DRAM %, L2 %, L1 %, plus DRAM bytes and shared-memory wavefronts per launch.
Usage: python profiles/summarize_ncu_step.py <ncu.csv> <out.json> [out.md]"""
import collections
import csv
import json
import sys

SHORT = {"gpu__time_duration.sum": "ns", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct_elapsed",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_active",
         "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
         "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
         "l1tex__data_bank_reads.sum": "bank_reads", "l1tex__data_bank_writes.sum": "bank_writes",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "sm__cycles_active.avg": "sm_cycles_active",
         "launch__registers_per_thread": "regs", "launch__shared_mem_per_block_dynamic": "dyn_smem"}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    col = {n: i for i, n in enumerate(rows[hdr])}
    d = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(col):
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        k = d.setdefault(int(r[0]), {"name": name, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
        m = SHORT.get(r[col["Metric Name"]])
        if m:
            v = float(r[col["Metric Value"]].replace(",", ""))
            k[m] = v * UNIT.get(r[col["Metric Unit"]], 1.0) if m in ("ns", "dram_rd", "dram_wr", "dyn_smem") else v
    return list(d.values())


def main():
    ks = load(sys.argv[1])
    marks = [i for i, k in enumerate(ks) if "sgd_kernel" in k["name"]]
    step = ks[marks[0] + 1:marks[1] + 1] if len(marks) >= 2 else ks
    total = sum(k.get("ns", 0.0) for k in step)
    groups = collections.OrderedDict()
    for k in step:
        key = "%s grid%s block%s" % (k["name"].split("<")[0] + ("<" + k["name"].split("<", 1)[1] if "<" in k["name"] and "conv_tc" in k["name"] else ""),
                                     k["grid"], k["block"])
        g = groups.setdefault(key, collections.defaultdict(float))
        g["launches"] += 1
        t = k.get("ns", 0.0)
        g["ns"] += t
        for m in ("tensor_pct_elapsed", "tensor_pct_active", "dram_pct", "l2_pct", "l1_pct", "sm_pct"):
            g[m] += k.get(m, 0.0) * t
        for m in ("dram_rd", "dram_wr", "smem_wavefronts", "bank_reads", "bank_writes"):
            g[m] += k.get(m, 0.0)
        g["regs"], g["dyn_smem"] = k.get("regs", 0.0), k.get("dyn_smem", 0.0)
    out = []
    for key, g in sorted(groups.items(), key=lambda kv: -kv[1]["ns"]):
        n, t = g["launches"], max(g["ns"], 1.0)
        out.append({"kernel": key, "launches": int(n), "time_us": g["ns"] / 1e3, "share_of_step": g["ns"] / total,
                    "tensor_pct_of_elapsed": g["tensor_pct_elapsed"] / t, "tensor_pct_of_active": g["tensor_pct_active"] / t,
                    "dram_pct": g["dram_pct"] / t, "l2_pct": g["l2_pct"] / t, "l1_pct": g["l1_pct"] / t, "sm_pct": g["sm_pct"] / t,
                    "dram_MB_per_launch": (g["dram_rd"] + g["dram_wr"]) / n / 1e6, "smem_lsu_wavefronts_per_launch": g["smem_wavefronts"] / n,
                    "regs": int(g["regs"]), "dyn_smem_KB": g["dyn_smem"] / 1024})
    doc = {"source": sys.argv[1], "launches_in_step": len(step), "step_time_us_serialised": total / 1e3, "kernels": out,
           "note": "ncu serialises launches and runs them cold: shares and percentages, not absolute times, compare with bench.py"}
    json.dump(doc, open(sys.argv[2], "w"), indent=1)
    lines = ["| kernel (grid, block) | launches | time (ncu) | share | tensor pipe % of elapsed (of active) | DRAM % | L2 % | L1/smem % | DRAM MB / launch | regs, dyn smem |", "|---|---|---|---|---|---|---|---|---|---|"]
    for r in out:
        if r["share_of_step"] < 0.004:
            continue
        lines.append("| `%s` | %d | %.0f us | %.1f %% | %.1f (%.1f) | %.1f | %.1f | %.1f | %.1f | %d, %.0f KB |" % (
            r["kernel"], r["launches"], r["time_us"], 100 * r["share_of_step"], r["tensor_pct_of_elapsed"], r["tensor_pct_of_active"],
            r["dram_pct"], r["l2_pct"], r["l1_pct"], r["dram_MB_per_launch"], r["regs"], r["dyn_smem_KB"]))
    md = "\n".join(lines)
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(md + "\n")
    print(md)


if __name__ == "__main__":
    main()
