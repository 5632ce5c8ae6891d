"""Opcode mix of one kernel from an `ncu --page source --csv` export: warp-instructions per unit and stall-sample share.
usage: python scripts/sass_mix.py <source.csv> <units (e.g. bands)>"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
ops, stall, tot = collections.Counter(), collections.Counter(), 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    src = r[idx["Source"]].strip()
    n = int(r[idx["Instructions Executed"]])
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = (m.group(2) if m else src).split(".")[0]
    ops[op] += n
    tot += n
    stall[op] += int(r[idx["# Samples"]])
ts = sum(stall.values()) or 1
print(f"total warp-instr/unit {tot / units:.1f}")
for op, n in ops.most_common(40):
    print(f"{op:12s} {n / units:8.2f}  stall-samples {100 * stall[op] / ts:5.1f}%")
