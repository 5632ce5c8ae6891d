"""Executed warp-instructions per source line of one kernel, from an ncu capture taken with --import-source on and the
object the kernel was built from (nvdisasm -g gives the line of every SASS instruction, the capture's source page
their executed counts).  usage: python scripts/sass_by_line.py <rep.ncu-rep> <obj.o> <kernel-substring e.g. ILi1E> <units> [top]
<units> = what to divide by (e.g. the bands the captured launch processed)."""
import collections, csv, os, re, subprocess, sys, tempfile
rep, obj, pat, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
want = {"ILi0E": "<(int)0>", "ILi1E": "<(int)1>", "ILi2E": "<(int)2>"}.get(pat, pat)
blk = next(b for b in blocks if want in b["name"])
hdr = blk["rows"][0]; ix = {h: i for i, h in enumerate(hdr)}
seen, ins = set(), []
for r in blk["rows"][1:]:
    a = r[ix["Address"]]
    if a in seen:
        continue
    seen.add(a)
    ins.append((int(a, 16), r[ix["Source"]].strip(), float(r[ix["Instructions Executed"]] or 0)))
ins.sort()
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if l.startswith(".text.") and pat in l)
end = next((i for i, l in enumerate(txt) if i > start and l.startswith("//------")), len(txt))
cur, lines = None, []
for l in txt[start:end]:
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2)))
assert len(lines) == len(ins), (len(lines), len(ins), "the object is not the build the capture was taken from")
agg = collections.Counter()
for (a, s, c), (off, line, t) in zip(ins, lines):
    agg[line] += c
tot = sum(agg.values())
print(f"{blk['name']}: {tot / units:.1f} executed warp-instructions per unit")
srcs = {}
for line, c in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    f, n = line
    if f not in srcs:
        p = next((os.path.join(d, f) for d, _, fs in os.walk(os.path.dirname(os.path.abspath(obj)) + "/../..") if f in fs), None)
        srcs[f] = open(p).read().split("\n") if p else []
    t = srcs[f][n - 1].strip()[:110] if srcs[f] else ""
    print(f"{c / units:7.2f}  {f}:{n}: {t}")
