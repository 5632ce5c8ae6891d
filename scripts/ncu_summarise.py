"""Turn ncu exports into the tracked summaries under profiles/.

  python scripts/ncu_summarise.py <tag> <raw.csv> <launches.csv> [--cells N --bands N --samples N]

<raw.csv>      = `ncu -i X.ncu-rep --page raw --csv` of a `--set full` capture (scripts/gpu_profile.sh)
<launches.csv> = `ncu --metrics gpu__time_duration.sum --csv` launch list of one bench run
Writes profiles/<tag>_ncu_summary.md, profiles/<tag>_launches.md and updates profiles/kernel_constants.json
(the measured per-unit constants bench.py's roofline uses)."""
import argparse, collections, csv, json, os, re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def short(name):
    m = re.search(r"(\w+_kernel)(<(?:\(int\))?(\d)>)?", name)
    if not m:
        return name[:40]
    k = m.group(1)
    if k == "align_kernel" and m.group(3):
        k += {"0": "", "1": "_fill_only", "2": "_backtrace_only"}[m.group(3)]
    return k


def to_bytes(v, unit):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag"); ap.add_argument("raw"); ap.add_argument("launches")
    ap.add_argument("--cells", type=float); ap.add_argument("--bands", type=float); ap.add_argument("--samples", type=float)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    out = os.path.join(ROOT, "profiles")
    consts_path = os.path.join(out, "kernel_constants.json")
    consts = json.load(open(consts_path)) if os.path.exists(consts_path) else {}

    rows = list(csv.reader(open(a.raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    md = [f"# ncu --set full summary, {a.tag}", "", a.note, ""]
    for r in rows[2:]:
        k = short(r[idx["Kernel Name"]])
        md += [f"## {k}   (`{r[idx['Kernel Name']][:90]}`)", "", "| metric | value | unit |", "|---|---|---|"]
        vals = {}
        for w in WANT:
            if w in idx:
                vals[w] = (r[idx[w]], units[idx[w]])
                md.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
        md.append("")
        if "dram__bytes_read.sum" in vals:
            tr = to_bytes(*vals["dram__bytes_read.sum"]) + to_bytes(*vals["dram__bytes_write.sum"])
            md.append(f"DRAM traffic per launch: {tr / 1e9:.3f} GB")
            if a.cells and k in ("align_kernel_fill_only", "banded_dp_kernel"):
                consts["align_dram_bytes_per_cell"] = tr / a.cells
                md.append(f"-> {tr / a.cells:.3f} B per DP cell (algorithmic floor 0.29 B/cell)")
            if a.samples and k == "seg_tile_kernel":
                consts["seg_tile_dram_bytes_per_sample"] = tr / a.samples
                md.append(f"-> {tr / a.samples:.2f} B per sample")
        if a.cells and k in ("align_kernel_fill_only", "banded_dp_kernel") and "smsp__thread_inst_executed.sum" in vals:
            ti = float(vals["smsp__thread_inst_executed.sum"][0])
            wi = float(vals["smsp__inst_executed.sum"][0])
            consts["align_thread_instr_per_cell"] = ti / a.cells
            consts["align_warp_instr_per_band"] = wi / a.bands if a.bands else None
            consts["align_issue_active_pct"] = float(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0])
            consts["source"] = f"profiles/{a.tag}_ncu_summary.md"
            md.append(f"-> {ti / a.cells:.1f} thread-instructions per DP cell, "
                      f"{wi / a.bands if a.bands else float('nan'):.1f} warp-instructions per band")
        md.append("")
    open(os.path.join(out, f"{a.tag}_ncu_summary.md"), "w").write("\n".join(md))

    # launch list: share of each kernel in the step
    rows = [r for r in csv.reader(open(a.launches)) if len(r) > 10 and r[0].isdigit()]
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for r in rows:
        k = short(r[4]); v = float(r[-1]); u = r[-2]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        tot[k] += ns; cnt[k] += 1
    total = sum(tot.values()) or 1
    md = [f"# ncu launch list, {a.tag}", "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none` over one `bench.py --reads 2000 --steps 1 --warmup 1` "
          "(value leg + e2e leg).  Cold-cache, serialised: compare SHARES with bench.py's stage_ms, not absolutes.", "",
          "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, ns in sorted(tot.items(), key=lambda kv: -kv[1]):
        md.append(f"| {k} | {cnt[k]} | {ns / 1e6:.2f} | {100 * ns / total:.1f} % |")
    open(os.path.join(out, f"{a.tag}_launches.md"), "w").write("\n".join(md) + "\n")
    json.dump(consts, open(consts_path, "w"), indent=1)
    print("wrote profiles/%s_*.md, kernel_constants.json:" % a.tag, consts)


if __name__ == "__main__":
    main()
