#!/usr/bin/env python
"""bench.py -- Msamples/s through event-detect + normalise + banded-align (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU code on the host cores

One step = one pass of the hot path over the whole workload: BASELINE.json configs[1], 100 000 synthetic R10.4.1
reads with a 30 kb N50 (~3*10^10 samples), per GPU (weak scaling: every rank gets its own length-balanced shard of a
N*100k-read batch, no collective on the data path).  `value` is measured with the inputs resident in HBM; `e2e`
goes through the public C-ABI call (dnb_submit) with host buffers, H2D and D2H inside the timed region: the signal is
DMA'd straight out of the page-locked workload buffer (no staging pass), queryToRef goes as CIGAR-like runs, the
results come back in the compact wire format (1 B + 4 B per event, 2 bits per alignment step).
`parity_check` runs a length-stratified subset of the SAME workload through the unmodified reference on the host and
compares bit for bit; a step whose reads did not all come back DNB_READ_OK fails the bench.
The value leg runs its device bins one at a time, so every launch has a clean CUDA-event duration for `roofline`
(align_kernel, issue-bound: thread-instructions per DP cell of the shipped build from profiles/kernel_constants.json x
cells / launch duration) and `roofline_segmentation` (HBM by contract + the issue view); the e2e leg keeps several
submissions in flight, three of them computing, whose kernels fill the ragged ends of each other's one-warp-per-read
alignment launches (--value-inflight 2 does the same for the resident leg; see config.value_leg).
Extra legs in the same line (N=1): `chain` (rows f1-f2: eventalign + DNN input tensors), `analogue` (configs[3]: LLR
calls per second), `ultra_long` (configs[2]); tuning-only options: --e2e-sweep, --value-inflight, --bin-samples.
"""
from __future__ import annotations

import argparse
import json
import os

# torch.distributed.run exports OMP_NUM_THREADS=1 to every rank unless the caller set it.  The host side of the e2e leg
# (dnb_submit's staging of ~78 GB per step into pinned memory) is an OpenMP loop, so one thread per rank would measure
# torchrun's default, not the library: give each rank its share of the host cores.  This has to happen before anything
# loads libgomp (numpy/torch/our library).
if "LOCAL_RANK" in os.environ and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    _local = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _local))
    # several submissions are in flight per rank, each from its own host thread with its own OpenMP team; idle teams
    # must sleep, not spin, when N ranks share the box's cores
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Msamples/s through event-detect+banded-align"
UNIT = "Msamples/s"
SAMPLES_PER_BASE = 12.5          # 5 kHz / 400 bp/s


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def pore_model():
    return np.load(os.path.join(ROOT, "tests", "golden", "pore_model_r10.4.1_400bps.npz"))["mean"].astype(np.float64)


def workload_lengths(n_reads_total: int, n50: float, seed: int) -> np.ndarray:
    from dnascent_b200 import synth
    return synth.lognormal_lengths(n_reads_total, n50, np.random.default_rng(seed))


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_sample_reads(n_reads: int, n50: float, seed: int):
    """Bounded sample of the same workload law, generated on the CPU (identical for both arms)."""
    from dnascent_b200 import synth
    mean = pore_model()
    lengths = workload_lengths(n_reads, n50, seed + 7919)
    ref = synth.make_reference(int(lengths.max()) + 50_000, seed + 1)
    return synth.simulate_batch(ref, lengths, mean, seed=seed + 2), ref, mean


def run_cpu_reference(reads, ref, mean, threads: int):
    """The reference's own read loop (detect.cpp:852-876 minus I/O and DNN).  oracle/_ref when it was built
    (unmodified reference sources), else the C port.  Returns (seconds, failed, kind)."""
    from oracle import refbind
    if refbind.available():
        R = refbind.Ref()
        R.set_model(refbind.PORE, mean, np.full(mean.size, 0.14))
        R.set_reference(ref)
        handles = [R.read_new(r) for r in reads]
        t, failed = R.bench_normalise(handles, threads)
        for h in handles:
            h.free()
        return t, failed, "reference"
    from oracle import portbind
    P = portbind.Port()
    t, failed = P.bench_normalise(reads, mean, threads)
    return t, failed, "port"


def run_cpu_chain(reads, ref, mean, threads: int):
    """The reference's normaliseEvents + eventalign + tensor builders per read (detect.cpp:876-888) on all host threads;
    only where oracle/_ref was built (the C port has no OpenMP loop for this stage)."""
    from oracle import refbind
    if not refbind.available():
        return None
    R = refbind.Ref()
    R.set_model(refbind.PORE, mean, np.full(mean.size, 0.14))
    R.set_reference(ref)
    handles = [R.read_new(r) for r in reads]
    t, failed = R.bench_chain(handles, threads)
    for h in handles:
        h.free()
    ns = sum(r.raw.size for r in reads)
    return {"value": ns / t / 1e6, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": f"{len(reads)} reads of the C2 length law ({ns} samples, {t:.1f} s wall, {failed} failed)"}


def synthetic_analogue_tables(mean: np.ndarray, seed: int = 99):
    """(unlabelled, analogue) Gaussian tables for the analogue bench leg: a seeded perturbation of the ONT table (T-containing
    9-mers shifted in the analogue one), same shape and value range as r10.4.1_{unlabelled,BrdU}_gaussian.model."""
    rng = np.random.default_rng(seed)
    n = mean.size
    unl = (mean + rng.normal(0, 0.02, size=n), rng.uniform(0.08, 0.20, size=n))
    idx = np.arange(n)
    has_t = np.zeros(n, dtype=bool)
    for j in range(9):
        has_t |= ((idx >> (2 * j)) & 3) == 1              # A=0 T=1 G=2 C=3 (src/data_IO.cpp:131-137)
    ana = (np.where(has_t, unl[0] + rng.normal(0, 0.15, size=n), 0.0), np.where(has_t, rng.uniform(0.10, 0.25, size=n), 0.0))
    return unl, ana


def run_cpu_hmm(mean, threads: int, seed: int):
    """The reference's detect --HMM loop body (normaliseEvents + llAcrossRead, detect.cpp:876-885) on all host threads,
    on a bounded sample of 10-kb reads; only where oracle/_ref was built."""
    from oracle import refbind
    from dnascent_b200 import synth
    if not refbind.available():
        return None
    R = refbind.Ref()
    R.set_model(refbind.PORE, mean, np.full(mean.size, 0.14))
    unl, ana = synthetic_analogue_tables(mean)
    R.set_model(refbind.UNLABELLED, *unl)
    R.set_model(refbind.ANALOGUE, *ana)
    ref = synth.make_reference(200_000, seed + 5)
    R.set_reference(ref)
    reads = synth.simulate_batch(ref, [10_000] * threads, mean, seed=seed + 6)
    handles = [R.read_new(r) for r in reads]
    t, failed, calls = R.bench_hmm(handles, threads, 12)
    for h in handles:
        h.free()
    return {"value": calls / t, "unit": "LLR calls/s (sites scored, both passes)", "cores": threads, "kind": "reference",
            "reads_per_s": len(reads) / t,
            "sample": f"{len(reads)} reads of 10 kb ({calls} calls, {t:.1f} s wall, {failed} failed)"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_reads = max(16 * cores, 64)
    reads, ref, mean = cpu_sample_reads(n_reads, args.n50, args.seed)
    n_samp = sum(r.raw.size for r in reads)
    for _ in range(args.warmup):
        run_cpu_reference(reads[: max(cores // 2, 2)], ref, mean, cores)
    t_tot, kind = 0.0, "reference"
    for _ in range(args.steps):
        t, _, kind = run_cpu_reference(reads, ref, mean, cores)
        t_tot += t
    v = n_samp * args.steps / t_tot / 1e6
    sample = f"{n_reads} reads of the C2 length law ({n_samp} samples) per step, all {cores} host threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: 100k synthetic reads, 30 kb N50 (bounded CPU sample)", "n50": args.n50},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ parity on the workload
def parity_check(ctx, W, mean, n_reads: int):
    """A length-stratified subset of the bench's own reads through the library (dnb_submit, the e2e path) and through the
    reference on the host (oracle/_ref: unmodified sources; else the C port), compared bit for bit: event boundaries,
    event means, alignment path, scalings, QC."""
    import types
    from oracle import refbind
    order = np.argsort(W.n_samples, kind="stable")
    pick = np.unique(order[np.linspace(0, order.size - 1, min(n_reads, order.size)).astype(np.int64)])
    b = ctx.submit_descs(W.descs(pick))
    ours = b.results()
    b.release()
    rd = [W.read(int(i)) for i in pick]
    t0 = time.time()
    if refbind.available():
        kind = "reference"
        R = refbind.Ref()
        R.set_model(refbind.PORE, mean, np.full(mean.size, 0.14))
        genome = b"".join(r["basecall"] for r in rd)             # each read is its own stretch of the "genome"
        R.set_reference(genome)
        handles, pos = [], 0
        for k, r in enumerate(rd):
            L = len(r["basecall"])
            handles.append(R.read_new(types.SimpleNamespace(
                name=f"p{k}", seq_bam=r["basecall"], flag=0, pos=pos, cigar=np.array([(L << 4) | 0], dtype=np.uint32), raw=r["raw"])))
            pos += L
        R.bench_normalise(handles, os.cpu_count() or 1)
        want = []
        for h in handles:
            o = h.outputs(staged=False)
            o["event_start"] = np.concatenate([[0], np.cumsum(o["event_raw_len"].astype(np.int64))])
            o["failed"] = o["align_event"].size == 0
            want.append(o)
            h.free()
    else:
        kind = "port"
        from oracle import portbind
        P = portbind.Port()
        want = []
        for r in rd:
            o = P.normalise(r["raw"], r["basecall"], r["refseq"], r["query_to_ref"], mean)
            o["failed"] = o["status"] != 0
            want.append(o)
    mism, detail = 0, []
    for k, (o, w) in enumerate(zip(ours, want)):
        ok = (o.status != 0) == bool(w["failed"])
        ok = ok and np.array_equal(o.event_mean.astype(np.float64), np.asarray(w["event_mean"], dtype=np.float64))
        ok = ok and np.array_equal(o.event_start.astype(np.int64), np.asarray(w["event_start"], dtype=np.int64))
        if ok and not w["failed"]:
            ok = (np.array_equal(o.eventAlignment[:, 0], w["align_event"]) and np.array_equal(o.eventAlignment[:, 1], w["align_kmer"])
                  and o.shift == w["shift"] and o.scale == w["scale"] and o.avg_log_emission == w["avg_log_emission"]
                  and o.spanned == w["spanned"] and o.maxGap == w["max_gap"])
        if not ok:
            mism += 1
            detail.append(int(pick[k]))
    return {"reads": int(pick.size), "mismatches": mism, "mismatching_reads": detail[:8], "against": kind,
            "samples": int(W.n_samples[pick].sum()), "longest_read_samples": int(W.n_samples[pick].max()),
            "compared": "event boundaries, event means, alignment path, shift/scale, avg_log_emission, spanned, max_gap (==)",
            "cpu_s": time.time() - t0}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
            }
            while not self.stop_flag:
                self.sm.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.2)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100_000, help="reads per GPU (configs[1] = 100 000)")
    ap.add_argument("--n50", type=float, default=30_000.0)
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--bin-samples", type=float, default=2.5e9, help="samples per device bin (value leg)")
    ap.add_argument("--e2e-bin-samples", type=float, default=1.0e9,
                    help="samples per dnb_submit call (e2e leg; B200 sweep at 100k reads, 3 computing at a time: 5e8 7 500, 8e8 8 280, "
                         "1.2e9 8 500 Msamples/s -- but 1.2e9 x 8 in flight peaks near the 180 GB of HBM and ran out of memory in one "
                         "of three runs, so the default stays below it)")
    ap.add_argument("--e2e-inflight", type=int, default=8,
                    help="dnb_submit calls in flight (B200 sweep, 60k reads: 4e8x4 74 %% of the resident value, 8e8x8 84 %%)")
    ap.add_argument("--value-inflight", type=int, default=1, help="value leg: resident bins run at a time (bins are cut at bin_samples / this)")
    ap.add_argument("--e2e-sweep", default="", help="tuning: extra e2e runs, comma-separated compute_slots:bin_samples:inflight[:interleave]")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pin", action="store_true", help="leave the workload pageable (e2e goes through pinned staging)")
    ap.add_argument("--extra-legs-all-ranks", action="store_true",
                    help="run the chain / ultra-long / analogue legs under torchrun too (default: only at --gpus 1; they are "
                         "extras next to the contract's line and their device-side data generation is not worth N x the time)")
    ap.add_argument("--ultra-reads", type=int, default=600,
                    help="reads per GPU of the configs[2] leg (lengths log-uniform in 100 kb - 1 Mb, inputs resident); 0 = skip")
    ap.add_argument("--analogue-reads", type=int, default=4000,
                    help="10-kb reads (whole job) of the configs[3] leg: normaliseEvents + llAcrossRead on the device (dnb_submit_llr); 0 = skip")
    ap.add_argument("--parity-reads", type=int, default=64, help="reads of the workload re-run on the CPU reference (0 = skip)")
    ap.add_argument("--chain-reads", type=int, default=4000,
                    help="reads (whole job) of the rows f1-f2 leg (dnb_submit_chain); 0 = skip")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from dnascent_b200 import _lib, api, bench_data, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; dnascent_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    peaks, peak_src = load_peaks()
    mean = pore_model()

    # ---- host memory: every rank keeps its shard as the loader would hand it over (int16 signal 2 B/sample + sequences),
    # page-locked in place; the compact results of the submissions in flight (~1.1 B/sample) come on top.  N ranks share
    # one box's RAM: shrink what is in flight first, the shard only as a last resort (and say so in the line).
    reads_per_gpu = args.reads
    try:
        import psutil
        avail = psutil.virtual_memory().available

        def inflight_bytes():
            return 1.5 * args.e2e_bin_samples * args.e2e_inflight * world
        while inflight_bytes() > 0.1 * avail and args.e2e_inflight > 4:
            args.e2e_inflight -= 1
        need = 1.1 * 2.1 * SAMPLES_PER_BASE * 25_100.0 * reads_per_gpu * world + inflight_bytes()   # mean read ~25.1 kb at N50 30 kb
        if need > 0.85 * avail:
            reads_per_gpu = max(int(reads_per_gpu * 0.85 * avail / need), 1000)
    except Exception:  # noqa: BLE001
        pass
    if world > 1:
        t = torch.tensor([reads_per_gpu], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        reads_per_gpu = int(t.item())
    reduced = reads_per_gpu != args.reads
    args.reads = reads_per_gpu

    # ---- workload: one N*reads batch, length-balanced across ranks (no data-path collective) ----
    lengths_all = workload_lengths(world * args.reads, args.n50, args.seed)
    mine = sharding.shard_reads(lengths_all, world)[rank]
    t0 = time.time()
    W = bench_data.generate(lengths_all[mine], mean, args.seed + 1000 * rank, device=f"cuda:{local}")
    gen_s = time.time() - t0
    n_samples = int(W.n_samples.sum())

    # the loader's buffers are page-locked (dnb_host_register): dnb_submit then DMAs every read from where it lies
    pin_s, pinned = 0.0, False
    if not args.no_pin:
        t0 = time.time()
        try:
            api.host_register(W.dac)
            api.host_register(W.seq)
            pinned = True
        except Exception as ex:  # noqa: BLE001 -- fall back to the staged path, visibly
            print(f"bench.py: dnb_host_register failed ({ex}); e2e leg uses pinned staging", file=sys.stderr)
        pin_s = time.time() - t0

    ctx = api.Context(device=local, result_format=api.RESULT_COMPACT)
    ctx.load_model(api.MODEL_PORE, mean)

    # ---- value leg: inputs resident in HBM ----
    # The alignment launch runs one warp per read, so the end of every launch is ragged (the last partial wave, the
    # longest chains) and leaves warp slots idle.  With --value-inflight C > 1 the resident bins are run by C host
    # threads (dnb_batch_run is thread-safe; every batch has its own stream), so that the kernels of one bin fill the
    # ragged end of another -- what dnb_submit's pipeline does end to end.  Overlapping launches have no clean duration,
    # so the stage split and the roofline then come from a second timed region that runs the same bins one at a time.
    conc = max(int(args.value_inflight), 1)
    bins = sharding.make_bins(W.n_samples, int(args.bin_samples / conc))
    batches = [ctx.upload_descs(W.descs(b)) for b in bins]
    stage_ms = {}
    counts = {}

    def run_one(b):
        b.run()
        ms, cnt = b.timings()
        for k, v in ms.items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
        for k, v in cnt.items():
            counts[k] = counts.get(k, 0) + v
        b.drop_workspace()

    def step(c=conc):
        if c <= 1:
            for b in batches:
                run_one(b)
        else:
            with ThreadPoolExecutor(max_workers=c) as ex:
                list(ex.map(run_one, batches))

    for _ in range(args.warmup):
        step()
    stage_ms.clear(); counts.clear()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    barrier()
    dt = time.perf_counter() - t0
    dt = max_over_ranks(dt)
    total_samples = sum_over_ranks(float(n_samples))
    value = total_samples * args.steps / dt / 1e6
    K = args.steps
    sequential = None
    if conc > 1:
        # second timed region: the same bins one at a time (clean per-launch CUDA-event durations)
        stage_ms.clear(); counts.clear()
        step(1)
        stage_ms.clear(); counts.clear()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(1)
        barrier()
        dt_seq = max_over_ranks(time.perf_counter() - t0)
        sequential = {"what": "the same resident bins run one at a time: the stage split and the roofline's launch durations come from here",
                      "value": total_samples * args.steps / dt_seq / 1e6, "ms_per_step": 1e3 * dt_seq / args.steps}
    sampler.stop_flag = True
    sampler.join(timeout=2)
    per_step = {k: v / K for k, v in stage_ms.items()}
    cnt_step = {k: v // K for k, v in counts.items()}
    for b in batches:
        b.release()
    # every read of the synthetic workload is an exact substring of the reference: all of them must align and pass QC
    ok_share = 1.0 - sum_over_ranks(float(cnt_step["failed_reads"])) / (world * args.reads)
    if ok_share < 0.999:
        raise SystemExit(f"bench.py: only {ok_share:.4%} of the reads came back DNB_READ_OK in the value leg")

    # ---- roofline of the dominant kernel (align_kernel: band fill + backtrace, one launch) and of segmentation ----
    clocks = sampler.summary()
    f_clk = (clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)) * 1e6
    prof = {}
    pj = os.path.join(ROOT, "profiles", "kernel_constants.json")
    if os.path.exists(pj):
        with open(pj) as f:
            prof = json.load(f)
    i_cell = prof.get("align_thread_instr_per_cell")           # whole launch (fill + backtrace) / DP cells, from ncu
    launch_s = (per_step["banded_dp"] + per_step["backtrace"]) / 1e3      # CUDA-event duration of the fused launch
    fill_s = per_step["banded_dp"] / 1e3                       # its band-fill share (in-kernel warp cycles)
    cells = cnt_step["cells"]
    issue_peak = 148 * 4 * 32 * f_clk                          # thread-instructions/s at 1 warp-instr/clk/scheduler
    roofline = {
        "kernel": "align_kernel<0> (band fill + backtrace, one fused launch; duration = CUDA events around the launch)",
        "bound": "issue",
        "achieved": (cells * i_cell / launch_s / 1e12) if i_cell else None, "peak": issue_peak / 1e12,
        "unit": "T thread-instr/s", "frac": (cells * i_cell / launch_s / issue_peak) if i_cell else None,
        "traffic": prof.get("align_dram_bytes_per_cell", None) and prof["align_dram_bytes_per_cell"] * cells,
        "constants_source": prof.get("source"), "constants_commit": prof.get("commit"),
        "launch_ms_per_step": 1e3 * launch_s, "cells_per_s": cells / launch_s, "thread_instr_per_cell": i_cell,
        "sm_clock_used_mhz": f_clk / 1e6,
        "issue_active_pct": prof.get("align_issue_active_pct"), "fp64_pipe_pct": prof.get("align_fp64_pipe_pct"),
        "xu_pipe_pct": prof.get("align_xu_pipe_pct"), "warps_active_pct": prof.get("align_warps_active_pct"),
        "fill_phase_view": {"what": "band fill only: launch duration x the fill's share of the in-kernel warp cycles",
                            "ms_per_step": 1e3 * fill_s, "cells_per_s": cells / fill_s},
        "hbm_view": {"bound": "hbm", "achieved": 0.29 * cells / launch_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": 0.29 * cells / launch_s / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes_per_cell": 0.29},
        "peak_source": peak_src,
    }
    seg_s = per_step["segmentation"] / 1e3
    seg_bytes = 2.0 * cnt_step["samples"] + 8.0 * cnt_step["events"]      # int16 DAC in, (u32 start, f32 mean) out
    i_samp = prof.get("seg_thread_instr_per_sample")
    roofline_seg = {"kernel": "seg_* (checkpoint/scan, tiles, stitch, events)", "bound": "hbm",
                    "achieved": seg_bytes / seg_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": seg_bytes / seg_s / 1e9 / peaks["hbm_gbs"],
                    "traffic": prof.get("seg_dram_bytes_per_sample") and prof["seg_dram_bytes_per_sample"] * cnt_step["samples"],
                    "algorithmic_bytes": "2 B/sample (int16 DAC) + 8 B/event", "peak_source": peak_src,
                    "ms_per_step": {"checkpoint_or_scan": per_step.get("seg_checkpoint"), "tiles": per_step.get("seg_tiles"),
                                    "stitch_events_redo": per_step.get("seg_rest")},
                    "issue_view": {"what": "the exact t-statistics make this kernel instruction-bound, not HBM-bound",
                                   "thread_instr_per_sample": i_samp,
                                   "frac": (cnt_step["samples"] * i_samp / seg_s / issue_peak) if i_samp else None}}

    # ---- e2e leg: host buffers through dnb_submit, H2D + D2H inside the timed region ----
    ctx.trim()          # the value leg's device blocks (inputs + one bin's workspace) sit idle in the library's cache: hand them back
    e2e_bins = sharding.make_bins(W.n_samples, int(args.e2e_bin_samples))
    descs = [W.descs(b) for b in e2e_bins]
    io_acc = [0, 0, 0]

    def one(d):
        b = ctx.submit_descs(d)
        b.wait()
        res0 = b.result(0)                      # touch a result: the step's outcome is read on the host
        _, cnt = b.timings()                    # counts[7]: reads of this submission whose status is not DNB_READ_OK
        bad = cnt["failed_reads"] + (res0.status != api.READ_OK)
        io = b.io_bytes()                       # counted by the library from the copies it made
        b.release()
        return io[0], io[1], bad

    def e2e_step():
        hb = db = nb = 0
        with ThreadPoolExecutor(max_workers=args.e2e_inflight) as ex:
            for h2d, d2h, bad in ex.map(one, descs):
                hb += h2d
                db += d2h
                nb += bad
        io_acc[0], io_acc[1], io_acc[2] = hb, db, nb

    for _ in range(min(args.warmup, 3)):
        e2e_step()
    api.host_stats(reset=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    dt_e = max_over_ranks(time.perf_counter() - t0)
    ph, ph_cnt = api.host_stats()
    host_phases = {"what": "thread-seconds per e2e step inside dnb_submit by phase (rank 0; summed over the submitting threads)",
                   "per_step_s": {k: round(v / args.steps, 4) for k, v in ph.items()}, "counters_timed_region": ph_cnt}
    e2e_value = total_samples * args.steps / dt_e / 1e6
    h2d_bytes, d2h_bytes, e2e_bad = io_acc
    if sum_over_ranks(float(e2e_bad)) > 0.001 * world * args.reads:
        raise SystemExit(f"bench.py: {e2e_bad} reads of an e2e step did not come back DNB_READ_OK")

    # ---- optional: the same e2e leg under other pipeline settings (tuning runs only; never part of the contract's line) ----
    e2e_sweep = None
    if args.e2e_sweep and world == 1:
        e2e_sweep = []
        ctx.trim()                      # the first context's cached device blocks would starve the second one
        for spec in args.e2e_sweep.split(","):
            slots, bsz, infl, order = (spec.split(":") + ["sorted"])[:4]
            os.environ["DNB_COMPUTE_SLOTS"] = slots
            c2 = api.Context(device=local, result_format=api.RESULT_COMPACT)
            c2.load_model(api.MODEL_PORE, mean)
            bins2 = sharding.make_bins(W.n_samples, int(float(bsz)))
            if order == "interleave":                       # longest, shortest, 2nd longest, 2nd shortest, ...
                k = len(bins2)
                bins2 = [bins2[i // 2] if i % 2 == 0 else bins2[k - 1 - i // 2] for i in range(k)]
            d2 = [W.descs(b) for b in bins2]

            def one2(d):
                b = c2.submit_descs(d)
                b.wait()
                b.result(0)
                b.release()

            def step2():
                with ThreadPoolExecutor(max_workers=int(infl)) as ex:
                    list(ex.map(one2, d2))
            step2()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(2):
                step2()
            torch.cuda.synchronize()
            dt2 = (time.perf_counter() - t0) / 2
            e2e_sweep.append({"compute_slots": int(slots), "bin_samples": float(bsz), "inflight": int(infl), "order": order,
                              "value": n_samples / dt2 / 1e6, "ms_per_step": 1e3 * dt2})
            print(f"bench.py: e2e sweep {e2e_sweep[-1]}", file=sys.stderr)
            c2.close()
        os.environ.pop("DNB_COMPUTE_SLOTS", None)

    # ---- parity check on the bench's own workload (BASELINE configs[1]: "throughput and bit-exactness check") ----
    parity = None
    if rank == 0 and args.parity_reads > 0:
        parity = parity_check(ctx, W, mean, args.parity_reads)
        if parity["mismatches"]:
            raise SystemExit(f"bench.py: parity check failed: {parity}")

    # ---- chain leg (SURVEY s.8 rows f1-f2): normaliseEvents -> eventalign -> DNN input tensors through dnb_submit_chain,
    # host buffers in, tensors out, on a bounded sample of this rank's shard (the tensors are 108 B per reference base)
    extras = world == 1 or args.extra_legs_all_ranks

    def guarded(leg):
        """An extra leg must never cost the contract's line: at --gpus 1 a failure is recorded in its place.  (Under
        torchrun the legs contain barriers, so there an exception still ends the run rather than deadlocking it.)"""
        if world > 1:
            return leg()
        try:
            return leg()
        except Exception as ex:  # noqa: BLE001
            return {"error": repr(ex)[:500]}

    chain = None

    def chain_leg():
        # N ranks share one box's host RAM (pinned results ~15 B/sample per submission in flight): the sample and
        # the submission size shrink with the number of ranks
        n_c = min(max(args.chain_reads // world, 200), W.n_reads)
        c_bin = 4.0e8 / world
        cidx = np.linspace(0, W.n_reads - 1, n_c).astype(np.int64)          # evenly through the shard: the same length law
        c_bins = [cidx[b] for b in sharding.make_bins(W.n_samples[cidx], int(c_bin))]
        c_descs = [W.descs(b) for b in c_bins]
        c_extras = []
        for d in c_descs:
            x = np.zeros(d.size, dtype=_lib.READ_EXTRA_DTYPE)
            x["ref_to_query"] = W.q2r.ctypes.data                            # exact `{L}M` reads: refToQuery is the identity
            x["ref_end"] = d["ref_len"]
            c_extras.append(x)
        acc = {"rows": 0, "h2d": 0, "d2h": 0, "ea_ms": 0.0, "ft_ms": 0.0, "bad": 0}

        def chain_one(k):
            b = ctx.submit_chain_descs(c_descs[k], c_extras[k], 50)
            t2 = b.stage2_timings()
            io = b.io_bytes()
            out = (b.feature_rows(), io[0] + t2["h2d_bytes"], io[1] + t2["d2h_bytes"], t2["eventalign_kernel_ms"],
                   t2["features_kernel_ms"])
            b.release()
            return out

        def chain_pass():
            for k in acc:
                acc[k] = 0
            with ThreadPoolExecutor(max_workers=3) as ex:
                for rows, hb, db, ea, ft in ex.map(chain_one, range(len(c_descs))):
                    acc["rows"] += rows; acc["h2d"] += hb; acc["d2h"] += db; acc["ea_ms"] += ea; acc["ft_ms"] += ft

        chain_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            chain_pass()
        barrier()
        dt_c = max_over_ranks(time.perf_counter() - t0) / 2
        c_samples = int(W.n_samples[cidx].sum())
        c_total = sum_over_ranks(float(c_samples))
        return {"what": "dnb_submit_chain: normaliseEvents -> eventalign -> DNN input tensors (rows f1-f2), host buffers in, "
                         f"tensors out, {c_bin:.1e}-sample submissions, 3 in flight",
                 "value": c_total / dt_c / 1e6, "unit": UNIT, "reads_per_gpu": int(n_c), "samples_per_gpu": c_samples,
                 "ms_per_pass": 1e3 * dt_c, "tensor_rows_per_gpu": int(acc["rows"]), "h2d_bytes_per_pass": int(acc["h2d"]),
                 "d2h_bytes_per_pass": int(acc["d2h"]), "eventalign_kernel_ms": acc["ea_ms"], "features_kernel_ms": acc["ft_ms"],
                 "eventalign_mode": "read-serial" if os.environ.get("DNB_EA_WINDOW_PARALLEL", "")[:1] == "0"
                                    else "window-parallel",
                 "cpu_reference": None}

    if args.chain_reads > 0 and extras:
        chain = guarded(chain_leg)

    # ---- ultra-long leg (BASELINE configs[2]): lengths log-uniform in [100 kb, 1 Mb], inputs resident in HBM.  One warp
    # walks one read's band chain, so a bin of few very long reads cannot fill the device: report how full it was
    # (warps against resident warp slots) and what the longest read's serial chain costs (the tail).
    ultra = None

    def ultra_leg():
        ctx.trim()                      # the library caches its device blocks; torch needs room to generate the reads
        torch.cuda.empty_cache()
        rng_u = np.random.default_rng(args.seed + 4242 + rank)
        lens_u = np.exp(rng_u.uniform(np.log(100_000), np.log(1_000_000), size=args.ultra_reads)).astype(np.int64)
        WU = bench_data.generate(lens_u, mean, args.seed + 4243 + 1000 * rank, device=f"cuda:{local}", ref_len=1_050_000)
        u_bins = sharding.make_bins(WU.n_samples, int(args.bin_samples))
        u_batches = [ctx.upload_descs(WU.descs(b)) for b in u_bins]
        u_ms, u_cnt, per_bin = {}, {}, []

        def ultra_step(record):
            for k, b in enumerate(u_batches):
                b.run()
                ms, cnt = b.timings()
                if record:
                    for k2, v2 in ms.items():
                        u_ms[k2] = u_ms.get(k2, 0.0) + v2
                    for k2, v2 in cnt.items():
                        u_cnt[k2] = u_cnt.get(k2, 0) + v2
                    if len(per_bin) < len(u_batches):
                        launch = ms["banded_dp"] + ms["backtrace"]
                        per_bin.append({"reads": int(u_bins[k].size), "samples": int(WU.n_samples[u_bins[k]].sum()),
                                        "longest_read_samples": int(WU.n_samples[u_bins[k]].max()),
                                        "warps": int(u_bins[k].size), "resident_warp_slots": 148 * 24,
                                        "occupancy_of_slots": u_bins[k].size / (148 * 24.0),
                                        "align_launch_ms": launch, "cells_per_s": cnt["cells"] / (launch / 1e3),
                                        "failed_reads": cnt["failed_reads"]})
                b.drop_workspace()

        ultra_step(False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            ultra_step(True)
        barrier()
        dt_u = max_over_ranks(time.perf_counter() - t0) / 2
        u_samples = sum_over_ranks(float(WU.n_samples.sum()))
        for b in u_batches:
            b.release()
        c2_cells_per_s = cells / launch_s
        for pb in per_bin:
            # the same cells at the rate the saturated C2 bins reach: what is above that is under-fill + tail
            pb["tail_factor"] = pb["cells_per_s"] and c2_cells_per_s / pb["cells_per_s"]
        return {"what": "configs[2]: lengths log-uniform in [100 kb, 1 Mb], inputs resident, dnb_batch_run per bin",
                 "value": u_samples / dt_u / 1e6, "unit": UNIT, "reads_per_gpu": int(args.ultra_reads),
                 "samples_per_gpu": int(WU.n_samples.sum()), "ms_per_pass": 1e3 * dt_u,
                 "stage_ms_per_pass": {k2: v2 / 2 for k2, v2 in u_ms.items()},
                 "failed_reads": int(u_cnt.get("failed_reads", 0)) // 2, "bins": per_bin,
                "saturated_cells_per_s_for_comparison": c2_cells_per_s}

    if args.ultra_reads > 0 and extras:
        ultra = guarded(ultra_leg)

    # ---- analogue leg (SURVEY s.8 row a15, BASELINE configs[3]): detect --HMM's loop body, normaliseEvents + llAcrossRead,
    # on 10-kb reads through dnb_submit_llr: host buffers in, (site, log-likelihoods) out; metric = LLR calls per second
    analogue = None

    def analogue_leg():
        ctx.trim()
        torch.cuda.empty_cache()
        n_a = max(args.analogue_reads // world, 100)
        unl, ana = synthetic_analogue_tables(mean)
        ctx.load_model(api.MODEL_UNLABELLED, *unl)
        ctx.load_model(api.MODEL_ANALOGUE, *ana)
        WA = bench_data.generate(np.full(n_a, 10_000), mean, args.seed + 77 + 1000 * rank, device=f"cuda:{local}")
        a_bins = sharding.make_bins(WA.n_samples, int(2.0e8))
        a_descs = [WA.descs(b) for b in a_bins]
        a_extras = []
        for dsc in a_descs:
            x = np.zeros(dsc.size, dtype=_lib.READ_EXTRA_DTYPE)
            x["ref_to_query"] = WA.q2r.ctypes.data
            x["ref_end"] = dsc["ref_len"]
            x["is_reverse"] = np.arange(dsc.size) & 1
            a_extras.append(x)
        a_acc = {}

        def analogue_one(k):
            b = ctx.submit_llr_descs(a_descs[k], a_extras[k], 12)
            tm = b.analogue_timings()
            io = b.io_bytes()
            b.release()
            return tm, io

        def analogue_pass():
            a_acc.clear()
            with ThreadPoolExecutor(max_workers=2) as ex:
                for tm, io in ex.map(analogue_one, range(len(a_descs))):
                    for k2, v2 in tm.items():
                        a_acc[k2] = a_acc.get(k2, 0) + v2
                    a_acc["h2d"] = a_acc.get("h2d", 0) + io[0]

        analogue_pass()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            analogue_pass()
        barrier()
        dt_a = max_over_ranks(time.perf_counter() - t0) / 2
        calls_total = sum_over_ranks(float(a_acc["calls"]))
        return {"what": "dnb_submit_llr: normaliseEvents -> llAcrossRead (T sites, event windows, both forward passes per site "
                            "on the device), host buffers in, per-site log-likelihoods out; 10-kb reads (configs[3] shape)",
                    "value": calls_total / dt_a, "unit": "LLR calls/s (sites scored, both passes)",
                    "reads_per_gpu": int(n_a), "reads_per_s": world * n_a / dt_a, "ms_per_pass": 1e3 * dt_a,
                    "candidate_sites_per_gpu": int(a_acc["candidate_sites"]), "calls_per_gpu": int(a_acc["calls"]),
                    "observations_per_gpu": int(a_acc["observations"]),
                    "forward_kernel_ms": a_acc["forward_kernel_ms"], "sites_kernel_ms": a_acc["sites_kernel_ms"],
                    "forward_kernel_calls_per_s": a_acc["calls"] / (a_acc["forward_kernel_ms"] / 1e3) if a_acc["forward_kernel_ms"] else None,
                    "d2h_bytes_per_pass": int(a_acc["d2h_bytes"]),
                    "tables": "synthetic (seeded perturbation of the ONT 9-mer table; the fitted BrdU/EdU tables are not "
                              "shipped to the GPU box -- parity on the real tables is tests/test_analogue_gpu.py)",
                "cpu_reference": None}

    if args.analogue_reads > 0 and extras:
        analogue = guarded(analogue_leg)

    # ---- CPU baseline (rank 0, N == 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        reads, ref, mean_c = cpu_sample_reads(max(16 * cores, 64), args.n50, args.seed)
        t, failed, kind = run_cpu_reference(reads, ref, mean_c, cores)
        ns = sum(r.raw.size for r in reads)
        cpu = {"value": ns / t / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{len(reads)} reads of the C2 length law ({ns} samples, {t:.1f} s wall, {failed} failed QC)"}
        try:    # SURVEY s.8(d): the same loop on ONE host thread, on a smaller sample of the same reads
            one = reads[:8]
            t1, _, _ = run_cpu_reference(one, ref, mean_c, 1)
            ns1 = sum(r.raw.size for r in one)
            cpu["single_thread"] = {"value": ns1 / t1 / 1e6, "unit": UNIT, "cores": 1,
                                    "sample": f"{len(one)} of those reads ({ns1} samples, {t1:.1f} s wall)"}
        except Exception as ex:  # noqa: BLE001 -- a reporting extra must never cost the bench line
            cpu["single_thread"] = {"error": repr(ex)}
        if chain is not None and "error" not in chain:
            try:
                chain["cpu_reference"] = run_cpu_chain(reads[: 2 * cores], ref, mean_c, cores)
            except Exception as ex:  # noqa: BLE001
                chain["cpu_reference"] = {"error": repr(ex)}
        if analogue is not None and "error" not in analogue:
            try:
                analogue["cpu_reference"] = run_cpu_hmm(mean_c, cores, args.seed)
            except Exception as ex:  # noqa: BLE001
                analogue["cpu_reference"] = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"configs[1]: {args.reads} synthetic R10.4.1 reads per GPU, N50 {int(args.n50)} b, "
                            f"int16 DAC input, generated on device (seed {args.seed})",
                "reads_per_gpu": args.reads, "samples_per_gpu": n_samples, "bins": len(bins),
                "reads_per_gpu_reduced_for_host_ram": reduced,
                "l2": "inputs (>= 60 GB per step) exceed the 126 MB L2; no flush needed",
                "parallelism": f"read-sharded x{world}, length-balanced, no collective",
                "value_inflight": conc,
                "value_leg": (f"{conc} resident device bins run at a time (one host thread each): the kernels of one fill the ragged end "
                              "of another's one-warp-per-read alignment launch; stage_ms_per_step and the roofline's launch durations "
                              "come from the sequential timed region (sequential_pass)" if conc > 1 else
                              "device bins run ONE AT A TIME (dnb_batch_run blocks), which gives every kernel launch a clean CUDA-event "
                              "duration for the roofline; the ragged end of each one-warp-per-read alignment launch is idle time here"),
                "sequential_pass": sequential,
                "reads_per_s": total_samples and (world * args.reads * args.steps / dt), "reads_ok_share": ok_share,
                "stage_ms_per_step": per_step, "counts_per_step": cnt_step, "generation_s": gen_s,
            },
            "roofline": roofline, "roofline_segmentation": roofline_seg, "cpu_baseline": cpu, "chain": chain,
            "analogue": analogue, "ultra_long": ultra, "parity_check": parity, "e2e_sweep": e2e_sweep,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": 1e3 * dt_e / args.steps, "inflight": args.e2e_inflight, "bins": len(e2e_bins),
                    "samples_per_submit": args.e2e_bin_samples, "failed_reads_per_step": int(e2e_bad),
                    "input": ("page-locked loader buffers, direct DMA per read" if pinned else "pageable, through pinned staging"),
                    "pipeline": "dnb_submit from several host threads: up to 3 submissions in their compute phase at a time, so the kernels "
                                "of one fill the ragged end of another's alignment launch (why e2e can exceed the sequential value leg)",
                    "result_format": "compact (u8 event lengths + f32 means, 2-bit alignment steps)",
                    "host_register_s": pin_s, "host_phases": host_phases,
                    "host_threads_per_rank": int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)),
                    "host_cores": os.cpu_count() or 1},
            "gpu_launches": int(cnt_step["launches"] * args.steps), "clocks": clocks,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
