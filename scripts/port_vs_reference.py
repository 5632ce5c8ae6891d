"""TEST INFRASTRUCTURE (CPU only): the plain-C oracle port against the unmodified reference (oracle/_ref) on the reads of
the statistical-parity workload (scripts/ea_statistical_parity.py: C2 length law, both strands, a third with 1 %
substitutions) -- every normaliseEvents output compared with ==.  The GPU tests compare the CUDA path with the port, so
this is the pin of the pin at scale.

    python scripts/port_vs_reference.py [n_reads] [max_len] > profiles/<tag>_port_vs_reference.json
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dnascent_b200 import synth
from oracle import portbind, refbind

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
max_len = int(sys.argv[2]) if len(sys.argv) > 2 else 80_000
mean = np.load("tests/golden/pore_model_r10.4.1_400bps.npz")["mean"].astype(np.float64)
rng = np.random.default_rng(4711)
lengths = np.clip(synth.lognormal_lengths(n, 30_000.0, rng), 1500, max_len)
ref = synth.make_reference(int(lengths.max()) + 100_000, 4712)
R = refbind.Ref()
R.set_model(refbind.PORE, mean, np.full(mean.size, 0.14))
R.set_reference(ref)
P = portbind.Port()
t0 = time.time()
bad, samples, failed = [], 0, 0
for i, L in enumerate(lengths):
    L = int(L)
    r = synth.simulate_read(ref, int(rng.integers(0, len(ref) - L)), L, bool(i & 1), mean, rng, name=f"s{i}",
                            sub_rate=0.01 if i % 3 == 0 else 0.0)
    samples += r.raw.size
    a = R.read_new(r).normalise(staged=True)
    b = P.normalise(r.raw, r.basecall, r.refseq, r.query_to_ref, mean)
    failed += a["align_event"].size == 0
    same = (np.array_equal(a["event_mean"], b["event_mean"]) and np.array_equal(a["align_event"], b["align_event"])
            and np.array_equal(a["align_kmer"], b["align_kmer"]) and a["shift"] == b["shift"] and a["scale"] == b["scale"]
            and a["rough_shift"] == b["rough_shift"] and a["rough_scale"] == b["rough_scale"]
            and a["avg_log_emission"] == b["avg_log_emission"] and a["spanned"] == b["spanned"] and a["max_gap"] == b["max_gap"]
            and np.array_equal(a["cleaned_signal"], b["cleaned_signal"]) and np.array_equal(a["cleaned_rank"], b["cleaned_rank"]))
    if not same:
        bad.append(i)
print(json.dumps({"what": "oracle port vs the unmodified reference, normaliseEvents outputs compared with == (events, alignment, rough and "
                          "refined scalings, cleaned vectors, QC scalars)", "reads": n, "samples": int(samples),
                  "reads_failed_qc_in_the_reference": int(failed), "reads_with_a_difference": len(bad), "first_bad_reads": bad[:10],
                  "wall_s": time.time() - t0}))
