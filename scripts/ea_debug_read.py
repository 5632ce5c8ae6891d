"""Debug helper: read #idx of scripts/ea_statistical_parity.py's workload through dnb_submit_chain (current eventalign
mode, see DNB_EA_WINDOW_PARALLEL) and through the unmodified reference; prints where the DNN input rows differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dnascent_b200 import api, synth
from oracle import refbind

n, max_len, idx = 2000, 80_000, int(sys.argv[1])
mean = np.load("tests/golden/pore_model_r10.4.1_400bps.npz")["mean"].astype(np.float64)
rng = np.random.default_rng(4711)
lengths = np.clip(synth.lognormal_lengths(n, 30_000.0, rng), 1500, max_len)
ref = synth.make_reference(int(lengths.max()) + 100_000, 4712)
r = None
for i, L in enumerate(lengths[: idx + 1]):
    L = int(L)
    r = synth.simulate_read(ref, int(rng.integers(0, len(ref) - L)), L, bool(i & 1), mean, rng, name=f"s{i}",
                            sub_rate=0.01 if i % 3 == 0 else 0.0)
R = refbind.Ref()
R.set_model(refbind.PORE, mean, np.full(mean.size, 0.14))
R.set_reference(ref)
h = R.read_new(r)
R.bench_chain([h], 1)
want = h.aligned_positions()
ctx = api.Context(0, result_format=api.RESULT_COMPACT)
ctx.load_model(api.MODEL_PORE, mean)
x = dict(ref_to_query=h.ref_to_query, is_reverse=h.is_reverse, ref_start=h.ref_start, ref_end=h.ref_end)
b = ctx.submit_chain([api.Read.from_synth(r, use_dac=True).with_runs()], [x], 50, want_records=True)
f = b.feature_results(want_records=True)[0]
o = b.results()[0]
b.release()
P = want["core"].size
print("mode", os.environ.get("DNB_EA_WINDOW_PARALLEL", "default(wp)"), "read", idx, "len", len(r.basecall), "rev", h.is_reverse,
      "rows ref", P, "ours", f["core"].size, "records", f["event"].size)
m = min(P, f["core"].size)
d = (np.any(f["signal"][:m] != want["signal"][:m], axis=1) | (f["core"][:m] != want["core"][:m]) | (f["coords"][:m] != want["coords"][:m])
     | (f["ref_index"][:m] != want["ref_index"][:m]) | (f["quality"][:m] != want["quality"][:m]))
bad = np.flatnonzero(d)
print("differing rows", bad.size, "first", bad[:5], "last", bad[-3:])
if bad.size:
    k = bad[0]
    print("ref   row", k, want["ref_index"][k], want["coords"][k], want["quality"][k], want["signal"][k][:6])
    print("ours  row", k, f["ref_index"][k], f["coords"][k], f["quality"][k], f["signal"][k][:6])
    only_q = (f["quality"][:m] != want["quality"][:m]) & ~(np.any(f["signal"][:m] != want["signal"][:m], axis=1))
    print("rows differing ONLY in quality (indelScore):", int(only_q.sum()))
