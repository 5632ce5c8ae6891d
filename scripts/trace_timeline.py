"""Summarise a DNB_TRACE_HOST=1 log: per-phase totals and GPU-idle estimate over a time window.
usage: python scripts/trace_timeline.py LOG [t0 t1]"""
import re, sys, collections
rx = re.compile(r"\[dnb host\] (\w+)\s+(.+?)\s+([\d.]+) ms\s+\(t=([\d.]+)\)")
ev = []
for ln in open(sys.argv[1], errors="ignore"):
    m = rx.search(ln)
    if m:
        who, what, ms, t = m.group(1), m.group(2).strip(), float(m.group(3)), float(m.group(4))
        ev.append((t - ms / 1e3, t, who, what))
t0 = float(sys.argv[2]) if len(sys.argv) > 2 else min(e[0] for e in ev)
t1 = float(sys.argv[3]) if len(sys.argv) > 3 else max(e[1] for e in ev)
tot = collections.defaultdict(float)
busy = []
for a, b, who, what in ev:
    if b < t0 or a > t1:
        continue
    tot[(who, what)] += b - a
    if what.startswith("phase"):
        busy.append((max(a, t0), min(b, t1)))
busy.sort()
u, cur = 0.0, None
for a, b in busy:
    if cur is None or a > cur[1]:
        if cur: u += cur[1] - cur[0]
        cur = [a, b]
    else:
        cur[1] = max(cur[1], b)
if cur: u += cur[1] - cur[0]
print(f"window {t0:.3f}..{t1:.3f} = {t1 - t0:.3f} s; some batch waiting on the GPU for {u:.3f} s ({100 * u / (t1 - t0):.0f} %)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"  {k[0]:7s} {k[1]:30s} {v:8.3f} s")
