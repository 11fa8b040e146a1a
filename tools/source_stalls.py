"""Per-role summary of an `ncu --set full --import-source on` capture of ONE kernel launch (tools/r2_step48.sh):
    ncu -i gpurun_out/s48_fprop.ncu-rep --page source --csv > /tmp/src.csv
    python tools/source_stalls.py /tmp/src.csv 0,0x1210,0x2a10,0x8650,0x10000 53760 > profiles/r2_source_stalls_<kernel>.txt
arguments: the CSV, the code offsets where the warp roles begin (prologue, producers, MMA issue, epilogue, end - read
off the role-dispatch branches in the SASS), and the number of (plane, issuing-warp) pairs = executions of one UTCHMMA.
Prints stall samples and executed warp-instructions per role, the stall reasons inside the MMA-issue role and the 40
most-sampled instructions."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, data = rows[1], rows[2:]
    ia, isrc, ismp, iex = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed"))
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    base = int(data[0][ia], 16)
    recs = []
    for r in data:
        st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
        recs.append((int(r[ia], 16) - base, r[isrc].strip(), int(r[ismp] or 0), int(r[iex] or 0), st, r))
    tot = sum(r[2] for r in recs)
    print(rows[0][1])
    print("stall samples %d over %d instructions" % (tot, len(recs)))
    if len(sys.argv) > 3:
        b = [int(x, 16) for x in sys.argv[2].split(",")]
        div = float(sys.argv[3])
        names = ["prologue", "producers", "mma issue", "epilogue / rest"]
        for i in range(len(b) - 1):
            ex = sum(r[3] for r in recs if b[i] <= r[0] < b[i + 1])
            sm = sum(r[2] for r in recs if b[i] <= r[0] < b[i + 1])
            print("%-16s [%#x, %#x): %9d warp-instructions executed (%.1f per UTCHMMA execution count), %5d samples = %4.1f %%"
                  % (names[i], b[i], b[i + 1], ex, ex / div, sm, 100.0 * sm / tot))
        agg = {}
        for r in recs:
            if b[2] <= r[0] < b[3]:
                for i, h in stall_cols:
                    agg[h] = agg.get(h, 0) + int(r[5][i] or 0)
        print("stall reasons inside the MMA-issue role:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    print("offset  samples  share  executed  instruction  [top stall reasons]")
    for off, src, smp, ex, st, _ in sorted(recs, key=lambda r: -r[2])[:40]:
        print("%#07x %6d %5.1f%% %9d  %-72s %s" % (off, smp, 100.0 * smp / tot, ex, src[:72], st))


if __name__ == "__main__":
    main()
