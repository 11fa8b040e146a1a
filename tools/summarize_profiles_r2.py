"""Round-2 evidence: turns the ncu CSV exports of tools/r2_step13.sh (gpurun_out/s13_*) into the tracked summaries
under profiles/.   python tools/summarize_profiles_r2.py [src-prefix] [tag]

  <tag>_launches_{train,infer}.csv            every launch of `ncu --metrics gpu__time_duration.sum --clock-control none`
  <tag>_launches_{train,infer}_by_kernel.csv  the same aggregated per kernel (share of the summed kernel time)
  <tag>_ncu_key_metrics_{train,infer}.json    `ncu --set full` over one whole training step / one sliding-window call:
                                              per launch duration, tensor-pipe %, SM %, L2 %, DRAM bytes, grid, regs
  <tag>_dram_traffic.json                     DRAM bytes per launch (read + write), averaged per kernel and phase
                                              (fwd/ = before the loss head, bwd/ = after; infer/) - bench.py's
                                              roofline.traffic
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")
PRE = sys.argv[1] if len(sys.argv) > 1 else "s13"
TAG = sys.argv[2] if len(sys.argv) > 2 else "r2"


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    m = re.match(r"(?:void )?([\w:]+(?:<[^(]*>)?)\(", name)
    return m.group(1) if m else name[:60]


def launches(which):
    path = os.path.join(SRC, "%s_launches_%s.csv" % (PRE, which))
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg, per = OrderedDict(), []
    for r in rows:
        k, ns = short(r[4]), float(r[-1])
        per.append((int(r[0]), k, r[7], r[8], ns))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(DST, "%s_launches_%s_by_kernel.csv" % (TAG, which)), "w") as f:
        f.write("kernel,launches,total_us,share,avg_us\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.4f,%.2f\n" % (k.replace(",", ";"), n, ns / 1e3, ns / total, ns / n / 1e3))
    with open(os.path.join(DST, "%s_launches_%s.csv" % (TAG, which)), "w") as f:
        f.write("id,kernel,block,grid,duration_ns\n")
        for p in per:
            f.write("%d,%s,\"%s\",\"%s\",%.0f\n" % (p[0], p[1].replace(",", ";"), p[2], p[3], p[4]))


WANT = OrderedDict([
    ("gpu__time_duration.sum", "duration_us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_throughput_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_throughput_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"), ("sm__cycles_elapsed.avg.per_second", "sm_ghz"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
])
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def full(which):
    path = os.path.join(SRC, "%s_full_%s_raw.csv" % (PRE, which))
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = OrderedDict(id=int(r[0]), kernel=short(r[hdr.index("Kernel Name")]))
        for w, nm in WANT.items():
            if w in hdr:
                i = hdr.index(w)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if nm.startswith("dram_r") or nm.startswith("dram_w"):
                    d[nm + "_bytes"] = v * SCALE.get(u, 1.0)
                elif nm == "duration_us":
                    d[nm] = v * SCALE.get(u, 1.0)
                else:
                    d[nm] = v
        out.append(d)
    return out


def main():
    os.makedirs(DST, exist_ok=True)
    traffic = {}
    for which in ("train", "infer"):
        launches(which)
        recs = full(which)
        json.dump(recs, open(os.path.join(DST, "%s_ncu_key_metrics_%s.json" % (TAG, which)), "w"), indent=0)
        if which == "train":
            heads = [i for i, d in enumerate(recs) if d["kernel"].startswith("head_fwd")]
            firsts = [i for i, d in enumerate(recs) if d["kernel"].startswith("conv3d_first_tc_kernel<0>")]
            adams = [i for i, d in enumerate(recs) if d["kernel"].startswith("adam_kernel")]
            # the first complete step of the capture (the 15-minute budget of the capture may cut the last one short)
            lo = firsts[0]
            hi = firsts[1] if len(firsts) > 1 else (adams[0] + 3 if adams and adams[0] > lo else len(recs))
            head = [h for h in heads if h > lo][0]
            for i, d in enumerate(recs[lo:hi], lo):
                phase = "fwd" if i <= head else "bwd"
                t = d.get("dram_read_bytes", 0.0) + d.get("dram_write_bytes", 0.0)
                traffic.setdefault("%s/%s" % (phase, d["kernel"]), []).append(t)
        else:
            for d in recs:
                t = d.get("dram_read_bytes", 0.0) + d.get("dram_write_bytes", 0.0)
                traffic.setdefault("infer/%s" % d["kernel"], []).append(t)
    json.dump(OrderedDict((k, dict(launches=len(v), avg_bytes_per_launch=sum(v) / len(v), max_bytes=max(v)))
                          for k, v in sorted(traffic.items())),
              open(os.path.join(DST, "%s_dram_traffic.json" % TAG), "w"), indent=1)
    print("wrote", sorted(f for f in os.listdir(DST) if f.startswith(TAG)))


if __name__ == "__main__":
    main()
