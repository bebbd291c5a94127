"""Coarse histogram of an `ncu --page source --csv` export: share of executed instructions and of
stall samples per block of SASS instructions (development aid)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 60
hdr = rows[1]
ia, isrc, iex, ith, ismp = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
data = []
for r in rows[2:]:
    try:
        data.append((r[ia], r[isrc], int(r[iex]), int(r[ith]), int(r[ismp])))
    except Exception:
        pass
tot = sum(d[2] for d in data); tots = sum(d[4] for d in data)
print("kernel:", rows[0][1][:90]); print("total warp-inst", tot, "samples", tots)
for i in range(0, len(data), chunk):
    seg = data[i:i + chunk]
    ex = sum(d[2] for d in seg); th = sum(d[3] for d in seg); sm = sum(d[4] for d in seg)
    if ex / tot < 0.004 and sm / tots < 0.004:
        continue
    ops = {}
    for d in seg:
        parts = d[1].split()
        op = parts[1] if parts[0].startswith('@') else parts[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda x: -x[1])[:4]
    print(f"{i:5d}: inst {ex/tot*100:5.1f}%  samples {sm/tots*100:5.1f}%  avg-threads {th/max(ex,1):5.1f}  {top}")
