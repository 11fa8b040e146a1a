"""Turns the ncu CSV exports under gpurun_out/ into the tracked summaries under profiles/.
    python tools/summarize_profiles.py r1"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")
os.makedirs(DST, exist_ok=True)


def short(name):
    m = re.search(r"::(\w+)\(", name) or re.search(r"(\w+)\(", name)
    return m.group(1) if m else name[:40]


def launches():
    path = os.path.join(SRC, "launches_%s.csv" % TAG)
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    per = []
    for r in rows:
        k, ns = short(r[4]), float(r[-1])
        per.append((int(r[0]), k, r[7], r[8], ns))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(DST, "%s_launches_by_kernel.csv" % TAG), "w") as f:
        f.write("kernel,launches,total_us,share,avg_us\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.4f,%.2f\n" % (k, n, ns / 1e3, ns / total, ns / n / 1e3))
    with open(os.path.join(DST, "%s_launches.csv" % TAG), "w") as f:
        f.write("id,kernel,block,grid,duration_ns\n")
        for p in per:
            f.write("%d,%s,%s,%s,%.0f\n" % p)
    return agg, total


WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "launch__shared_mem_per_block_dynamic"]


def raw(name):
    path = os.path.join(SRC, "prof_%s_%s_raw.csv" % (TAG, name))
    if not os.path.exists(path):
        return []
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = OrderedDict()
        d["kernel"] = short(r[hdr.index("Kernel Name")])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = "%s %s" % (r[i], units[i])
        out.append(d)
    return out


def main():
    launches()
    caps = OrderedDict((n, raw(n)) for n in ("fprop", "dgrad", "wgrad", "bw"))
    json.dump(caps, open(os.path.join(DST, "%s_ncu_key_metrics.json" % TAG), "w"), indent=1)
    traffic = {}
    for n, recs in caps.items():
        for d in recs:
            try:
                def num(s):
                    v, u = s.split()[0], (s.split() + [""])[1]
                    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
                    return float(v.replace(",", "")) * scale
                t = num(d["dram__bytes_read.sum"]) + num(d["dram__bytes_write.sum"])
                traffic.setdefault("%s/%s" % (n, d["kernel"]), []).append(t)
            except Exception:
                pass
    json.dump({k: sum(v) / len(v) for k, v in traffic.items()},
              open(os.path.join(DST, "%s_dram_traffic.json" % TAG), "w"), indent=1)
    print("wrote", sorted(os.listdir(DST)))


if __name__ == "__main__":
    main()
