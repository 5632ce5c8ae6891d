"""Turn ncu exports into the tracked summaries under profiles/ and the per-unit constants bench.py's roofline uses.

  python scripts/ncu_summarise.py <tag> --raw a_raw.csv [b_raw.csv ...] --bench-log run.log [--llr-log run.log] [--note "..."]

<x_raw.csv>  = `ncu -i X.ncu-rep --page raw --csv` of a `--set full --metrics smsp__thread_inst_executed.sum` capture
<run.log>    = stdout of the bench.py run the capture was taken from (its JSON line gives the units the captured
               launches processed: samples, events, bands, DP cells of the step; the capture must be of a run whose
               workload is ONE device bin so that a launch == the step)
Writes profiles/<tag>_ncu_summary.md and profiles/kernel_constants.json (with the git commit the kernels were built
from, so that the constants can be tied to the shipped build)."""
import argparse, csv, json, os, re, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def short(name):
    m = re.search(r"(\w+_kernel)(<([^>]*)>)?", name)
    if not m:
        return name[:40]
    k = m.group(1)
    if k == "align_kernel" and m.group(3):
        k += {"0": " (fused fill + backtrace)", "1": " (fill only, DNB_SPLIT_ALIGN=1)", "2": " (backtrace only, DNB_SPLIT_ALIGN=1)"}.get(
            m.group(3).replace("(int)", ""), "")
    return k


def to_bytes(v, unit):
    try:
        f = float(v)
    except ValueError:
        return float("nan")
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def fnum(v):
    try:
        return float(v)
    except ValueError:
        return float("nan")


def bench_line(path):
    for ln in reversed(open(path).read().splitlines()):
        ln = ln.strip()
        if ln.startswith("{") and '"counts_per_step"' in ln:
            return json.loads(ln)
    raise SystemExit(f"no bench JSON line in {path}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--raw", nargs="+", required=True)
    ap.add_argument("--bench-log", required=True)
    ap.add_argument("--llr-log")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    out = os.path.join(ROOT, "profiles")
    line = bench_line(a.bench_log)
    cnt = line["config"]["counts_per_step"]
    assert line["config"]["bins"] == 1, "the captured run must have one device bin (a launch == the step)"
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    dirty = bool(subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", "dnascent_b200/csrc"], capture_output=True, text=True).stdout.strip())
    consts = {"source": f"profiles/{a.tag}_ncu_summary.md", "commit": commit + ("+uncommitted csrc changes" if dirty else ""),
              "captured_step": {k: cnt[k] for k in ("samples", "events", "bands", "cells")}}
    md = [f"# ncu --set full summary, {a.tag}", "", a.note, "",
          f"Kernels built from commit `{consts['commit']}`.  Captured step: {cnt['samples']:.3e} samples, {cnt['events']:.3e} events, "
          f"{cnt['bands']:.3e} bands, {cnt['cells']:.3e} DP cells (one device bin, so one launch of each kernel == the step).  "
          "Durations under ncu are cold-cache and serialised: not bench values.", ""]
    llr_sites = None
    if a.llr_log:
        ll = bench_line(a.llr_log)
        llr_sites = ll["analogue"]["calls_per_gpu"]
    for raw in a.raw:
        rows = list(csv.reader(open(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            name = r[idx["Kernel Name"]]
            k = short(name)
            md += [f"## {k}   (`{name[:100]}`)", "", "| metric | value | unit |", "|---|---|---|"]
            vals = {}
            for w in WANT:
                if w in idx:
                    vals[w] = (r[idx[w]], units[idx[w]])
                    md.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
            stalls = sorted(((fnum(r[i]), STALL.match(h).group(1)) for h, i in idx.items() if STALL.match(h)), reverse=True)
            md += ["", "top warp stall reasons (per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6] if v == v), ""]
            ti = fnum(vals.get("smsp__thread_inst_executed.sum", ("nan",))[0])
            wi = fnum(vals.get("smsp__inst_executed.sum", ("nan",))[0])
            dram = to_bytes(*vals.get("dram__bytes_read.sum", ("nan", ""))) + to_bytes(*vals.get("dram__bytes_write.sum", ("nan", "")))
            pct = lambda key: fnum(vals.get(key, ("nan",))[0])
            if "align_kernel" in name and "0>" in name.replace("(int)", ""):
                consts.update(align_thread_instr_per_cell=ti / cnt["cells"], align_warp_instr_per_band=wi / cnt["bands"],
                              align_dram_bytes_per_cell=dram / cnt["cells"],
                              align_issue_active_pct=pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                              align_fp64_pipe_pct=pct("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                              align_xu_pipe_pct=pct("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                              align_warps_active_pct=pct("sm__warps_active.avg.pct_of_peak_sustained_active"))
                md += [f"per unit: **{ti / cnt['cells']:.2f} thread-instructions per DP cell**, {wi / cnt['bands']:.1f} warp-instructions per band, "
                       f"{dram / cnt['cells']:.3f} DRAM bytes per cell (algorithmic floor 0.29 for the fill alone)", ""]
            elif "align_kernel" in name and "1>" in name.replace("(int)", ""):
                consts.update(fill_thread_instr_per_cell=ti / cnt["cells"], fill_dram_bytes_per_cell=dram / cnt["cells"])
                md += [f"per unit: {ti / cnt['cells']:.2f} thread-instructions per DP cell, {wi / cnt['bands']:.1f} warp-instructions per band, "
                       f"{dram / cnt['cells']:.3f} DRAM bytes per cell (algorithmic floor 0.29)", ""]
            elif "seg_tile_kernel" in name:
                consts.update(seg_tile_thread_instr_per_sample=ti / cnt["samples"], seg_tile_dram_bytes_per_sample=dram / cnt["samples"])
                md += [f"per unit: {ti / cnt['samples']:.1f} thread-instructions per sample, {dram / cnt['samples']:.2f} DRAM bytes per sample "
                       "(algorithmic 2 B/sample in + 8 B/event out = 3.5)", ""]
            elif "seg_scan_kernel" in name:
                consts.update(seg_scan_thread_instr_per_sample=ti / cnt["samples"], seg_scan_dram_bytes_per_sample=dram / cnt["samples"])
                md += [f"per unit: {ti / cnt['samples']:.1f} thread-instructions per sample, {dram / cnt['samples']:.2f} DRAM bytes per sample", ""]
            elif "theil_sen_kernel" in name:
                n_reads = fnum(vals["launch__grid_size"][0])
                consts.update(theil_sen_thread_instr_per_read=ti / n_reads)
                md += [f"per unit: {ti / n_reads / 1e6:.1f} M thread-instructions per read ({n_reads:.0f} reads)", ""]
            elif "llr_forward_kernel" in name and llr_sites:
                consts.update(llr_thread_instr_per_site=ti / llr_sites,
                              llr_fp64_pipe_pct=pct("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                              llr_issue_active_pct=pct("smsp__issue_active.avg.pct_of_peak_sustained_active"))
                md += [f"per unit: {ti / llr_sites / 1e6:.2f} M thread-instructions per site ({llr_sites} sites, both passes)", ""]
    if "seg_tile_thread_instr_per_sample" in consts and "seg_scan_thread_instr_per_sample" in consts:
        consts["seg_thread_instr_per_sample"] = consts["seg_tile_thread_instr_per_sample"] + consts["seg_scan_thread_instr_per_sample"]
        consts["seg_dram_bytes_per_sample"] = consts["seg_tile_dram_bytes_per_sample"] + consts["seg_scan_dram_bytes_per_sample"]
    with open(os.path.join(out, f"{a.tag}_ncu_summary.md"), "w") as f:
        f.write("\n".join(md) + "\n")
    # a metric ncu could not collect (it prints -nan when a replay pass fails) becomes null, not the invalid JSON token NaN
    consts = {k: (None if isinstance(v, float) and v != v else v) for k, v in consts.items()}
    with open(os.path.join(out, "kernel_constants.json"), "w") as f:
        json.dump(consts, f, indent=1, allow_nan=False)
    print(json.dumps(consts, indent=1))


if __name__ == "__main__":
    main()
