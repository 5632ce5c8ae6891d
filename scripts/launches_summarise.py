"""ncu launch list (`--metrics gpu__time_duration.sum --csv`) -> profiles/<tag>_launches.md: per kernel the number of
launches, the summed duration and its share.  Durations under ncu are cold-cache and serialised: compare shares only.
usage: python scripts/launches_summarise.py <tag> <launches.csv> [note]"""
import collections, csv, os, re, sys
tag, path = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
acc = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
    unit, val = r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ms
tot = sum(v[1] for v in acc.values())
OURS = re.compile(r"^(seg_|align_kernel|wp_|features_kernel|theil_sen|quantile_kernel|scale_events|compact_|count_long|ranks_kernel|"
                  r"expand_q2r|zero_padding|llr_|hmm_|eventalign_kernel)")
ours = [k for k in acc if OURS.match(k)]
lines = [f"# launch list, {tag}", "", note, "",
         f"{sum(v[0] for v in acc.values())} launches, {tot:.1f} ms under ncu in total; library kernels of this repo: "
         f"{sum(acc[k][0] for k in ours)} launches, {sum(acc[k][1] for k in ours):.1f} ms.  torch kernels in the list belong to the synthetic-input generation (bench_data.py), which is outside every timed region.",
         "", "| kernel | launches | ms (sum) | share of our kernels |", "|---|---|---|---|"]
t_ours = sum(acc[k][1] for k in ours) or 1.0
for k in sorted(ours, key=lambda k: -acc[k][1]):
    lines.append(f"| `{k}` | {acc[k][0]} | {acc[k][1]:.2f} | {100 * acc[k][1] / t_ours:.1f} % |")
open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"{tag}_launches.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:30]))
