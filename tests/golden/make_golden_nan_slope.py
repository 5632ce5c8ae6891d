"""Fixtures for the two reads of the round-2 statistical parity workload (scripts/ea_statistical_parity.py, 2000 reads)
whose Theil-Sen slopes contain a NaN (0/0: two cleaned points with identical signal and model level).  std::sort with a
NaN in the range is outside its contract, so the reference's median depends on where libstdc++'s introsort happens to
leave the NaN: AFTER element [size/2] on read 1797, BEFORE it on read 1803.  The expected values come from the
UNMODIFIED reference (oracle/_ref); run in the build container:

    python tests/golden/make_golden_nan_slope.py      # writes read_theilsen_nan_slope.npz (1797) and ..._b.npz (1803)
"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np
from dnascent_b200 import synth
from oracle import refbind

WANT = {1797: "read_theilsen_nan_slope.npz", 1803: "read_theilsen_nan_slope_b.npz"}
n, max_len = 2000, 80_000
mean = np.load(os.path.join(HERE, "pore_model_r10.4.1_400bps.npz"))["mean"].astype(np.float64)
rng = np.random.default_rng(4711)
lengths = np.clip(synth.lognormal_lengths(n, 30_000.0, rng), 1500, max_len)
ref = synth.make_reference(int(lengths.max()) + 100_000, 4712)
R = refbind.Ref()
R.set_model(refbind.PORE, mean, np.full(mean.size, 0.14))
R.set_reference(ref)
for i, L in enumerate(lengths[: max(WANT) + 1]):
    L = int(L)
    r = synth.simulate_read(ref, int(rng.integers(0, len(ref) - L)), L, bool(i & 1), mean, rng, name=f"s{i}",
                            sub_rate=0.01 if i % 3 == 0 else 0.0)
    if i not in WANT:
        continue
    o = R.read_new(r).normalise(staged=True)
    # where did the NaN land?  (for the record printed below)
    sig, rk = o["cleaned_signal"], o["cleaned_rank"]
    eff = sig.size - 100
    skip = max(eff // 1000, 1) if eff > 1000 else 1
    npnt = min(eff, 1000)
    idx = 50 + np.arange(npnt) * skip
    x = (sig[idx] - o["rough_shift"]) / o["rough_scale"]
    y = mean[rk[idx]]
    iu = np.triu_indices(npnt, 1)
    with np.errstate(all="ignore"):
        sl = (y[iu[0]] - y[iu[1]]) / (x[iu[0]] - x[iu[1]])
    srt = np.sort(sl[~np.isnan(sl)])
    m = sl.size // 2
    slope_ref = -1.0 / ((o["shift"] - o["rough_shift"]) / o["rough_scale"]) if False else None
    print(f"read {i}: {int(np.isnan(sl).sum())} NaN slope(s) of {sl.size}; NaN-last median {srt[m]!r}, NaN-first median {srt[m - 1]!r}; "
          f"reference scale/rough_scale = {o['scale'] / o['rough_scale']!r} (1/median: last {1 / srt[m]!r}, first {1 / srt[m - 1]!r})")
    np.savez_compressed(os.path.join(HERE, WANT[i]), dac=r.dac, basecall=np.frombuffer(r.basecall, dtype=np.uint8),
                        refseq=np.frombuffer(r.refseq, dtype=np.uint8), q2r=r.query_to_ref,
                        shift=o["shift"], scale=o["scale"], rough_shift=o["rough_shift"], rough_scale=o["rough_scale"],
                        n_align=np.int64(o["align_event"].size))
