"""Row f1/f2 statistical parity: the window-parallel eventalign's emission arithmetic is not the
reference's literal expression (eventalign_core.cuh: log c + y with a reciprocal multiply instead of log(c * exp(y))), so
a Viterbi path could in principle flip where two candidates tie to ~1e-15.  This run bounds that empirically: N reads of
the C2 length law (both strands, a third with 1 % substitutions) go through dnb_submit_chain on the GPU and through the
UNMODIFIED reference's normaliseEvents + eventalign + tensor builders (oracle/_ref, all host cores), and every DNN
input row (20 scaled samples, core / residual index, coordinate) -- i.e. which event was assigned to which reference
position in which state -- is compared bit for bit.

    python scripts/ea_statistical_parity.py [n_reads] [max_len] > gpurun_out/ea_statistical_parity.json
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dnascent_b200 import api, synth
from oracle import refbind

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
max_len = int(sys.argv[2]) if len(sys.argv) > 2 else 80_000
mean = np.load("tests/golden/pore_model_r10.4.1_400bps.npz")["mean"].astype(np.float64)
rng = np.random.default_rng(4711)
lengths = np.clip(synth.lognormal_lengths(n, 30_000.0, rng), 1500, max_len)
ref = synth.make_reference(int(lengths.max()) + 100_000, 4712)
t0 = time.time()
reads = []
for i, L in enumerate(lengths):
    L = int(L)
    reads.append(synth.simulate_read(ref, int(rng.integers(0, len(ref) - L)), L, bool(i & 1), mean, rng, name=f"s{i}",
                                     sub_rate=0.01 if i % 3 == 0 else 0.0))
gen_s = time.time() - t0

# ---- reference: normaliseEvents + eventalign + tensors, all cores ----
R = refbind.Ref()
R.set_model(refbind.PORE, mean, np.full(mean.size, 0.14))
R.set_reference(ref)
handles = [R.read_new(r) for r in reads]
cores = os.cpu_count() or 1
t_ref, failed_ref = R.bench_chain(handles, cores)

# ---- ours: one resident chain ----
ctx = api.Context(0, result_format=api.RESULT_COMPACT)
ctx.load_model(api.MODEL_PORE, mean)
extra = []
for r, h in zip(reads, handles):
    extra.append(dict(ref_to_query=h.ref_to_query, is_reverse=h.is_reverse, ref_start=h.ref_start, ref_end=h.ref_end))
t0 = time.time()
feats, norm = [], []
step = 250
for k in range(0, n, step):
    b = ctx.submit_chain([api.Read.from_synth(r, use_dac=True).with_runs() for r in reads[k:k + step]], extra[k:k + step], 50)
    norm += b.results()
    feats += b.feature_results()
    b.release()
t_gpu = time.time() - t0

rows = diff_rows = diff_reads = failed_both = status_mismatch = events = 0
bad = []
for i, (h, f, o) in enumerate(zip(handles, feats, norm)):
    events += int(o.event_mean.size)
    want = h.aligned_positions()
    ref_failed = want["core"].size == 0
    ours_failed = f["status"] != 0 or f["core"].size == 0
    if ref_failed or ours_failed:
        failed_both += ref_failed and ours_failed
        if ref_failed != ours_failed:
            status_mismatch += 1
            bad.append(i)
        continue
    P = want["core"].size
    rows += P
    if f["core"].size != P:
        diff_reads += 1; diff_rows += abs(int(f["core"].size) - P); bad.append(i)
        continue
    d = (np.any(f["signal"] != want["signal"], axis=1) | (f["core"] != want["core"]) | (f["residual"] != want["residual"])
         | (f["coords"] != want["coords"]) | (f["ref_index"] != want["ref_index"]) | (f["query_index"] != want["query_index"])
         | (f["quality"] != want["quality"]))
    if d.any():
        diff_reads += 1; diff_rows += int(d.sum()); bad.append(i)
    h.free()
print(json.dumps({
    "what": "dnb_submit_chain (window-parallel eventalign + tensors) vs the unmodified reference's normaliseEvents + eventalign + "
            "make*Tensor, every DNN input row compared with ==",
    "reads": n, "bases": int(lengths.sum()), "events_viterbi_steps": events, "tensor_rows_compared": rows,
    "reads_with_a_difference": diff_reads, "differing_rows": diff_rows, "status_mismatches": status_mismatch,
    "reads_failed_in_both": int(failed_both), "first_bad_reads": bad[:8],
    "reference_wall_s": t_ref, "reference_cores": cores, "gpu_wall_s_incl_python": t_gpu, "generation_s": gen_s}))
