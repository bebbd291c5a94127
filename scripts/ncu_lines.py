"""Per-source-line shares of an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` export:
executed warp instructions and stall samples attributed to each CUDA-C line (development aid)."""
import collections
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pytorch3d_pointops_b200", "csrc")
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ia, iex, ism, ith = (hdr.index(k) for k in ("Address", "Instructions Executed", "# Samples", "Thread Instructions Executed"))
        continue
    if hdr is None or len(r) <= max(ia, iex, ism, ith) or r[ia] == "":
        continue
    try:
        ln, ex, sm, th = int(r[0]), int(r[iex]), int(r[ism]), int(r[ith])
    except ValueError:
        continue
    a = agg[(cur_file, ln)]
    a[0] += ex
    a[1] += sm
    a[2] += th
    tot[0] += ex
    tot[1] += sm
print("total warp-inst", tot[0], "samples", tot[1])
srcs = {}


def src(f, ln):
    if f not in srcs:
        p = os.path.join(csrc, f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return srcs[f][ln - 1].strip()[:100] if 0 < ln <= len(srcs[f]) else ""


byfile = collections.defaultdict(lambda: [0, 0])
for (f, ln), (ex, sm, th) in agg.items():
    byfile[f][0] += ex
    byfile[f][1] += sm
for f, (ex, sm) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"{f}: inst {ex / tot[0] * 100:5.1f}%  samples {sm / tot[1] * 100:5.1f}%")
for (f, ln), (ex, sm, th) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:4d} inst {ex / tot[0] * 100:5.2f}% smp {sm / tot[1] * 100:5.2f}% thr {th / max(ex, 1):4.1f} | {src(f, ln)}")
