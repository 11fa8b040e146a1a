"""Opcode histogram per kernel of the shipped library (`cuobjdump -sass libfetalb200.so`): the committed evidence that
the conv kernels issue tcgen05 (UTCHMMA / UTCBAR / LDTM / STTM) and TMA (UTMALDG / UTMASTG / UBLKCP) instructions.
    python tools/sass_histogram.py > profiles/r2_sass_histogram.json"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fetal-mri-segmentation_b200", "fetal_net", "libfetalb200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS",
       "HMMA", "FFMA", "HFMA2", "DFMA", "DADD", "LDG", "STG", "LDS", "STS", "RED", "ATOM", "ATOMG", "SHFL", "BAR",
       "MUFU", "ACQBULK", "ELECT", "F2FP", "UGETNEXTWORKID")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = demangle(list(kernels))
    res = collections.OrderedDict()
    for k, c in kernels.items():
        full = names.get(k, k).replace("(anonymous namespace)::", "")
        m = re.match(r"(?:void )?([\w:]+(?:<[^(]*>)?)\(", full)
        short = m.group(1) if m else full[:80]
        d = collections.OrderedDict(total=sum(c.values()))
        d.update((op, c[op]) for op in KEY if c.get(op))
        d["top"] = ", ".join("%s %d" % kv for kv in c.most_common(8))
        # template instantiations of the same kernel: keep every one, keyed by the demangled name
        key, i = short, 1
        while key in res:
            i += 1
            key = "%s #%d" % (short, i)
        res[key] = d
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
