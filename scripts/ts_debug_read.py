"""Debug helper: read #idx of the statistical-parity workload: normaliseEvents on the GPU vs the CPU port, scalings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dnascent_b200 import api, synth
from oracle import portbind
n, max_len, idx = 2000, 80_000, int(sys.argv[1])
mean = np.load("tests/golden/pore_model_r10.4.1_400bps.npz")["mean"].astype(np.float64)
rng = np.random.default_rng(4711)
lengths = np.clip(synth.lognormal_lengths(n, 30_000.0, rng), 1500, max_len)
ref = synth.make_reference(int(lengths.max()) + 100_000, 4712)
for i, L in enumerate(lengths[: idx + 1]):
    L = int(L)
    r = synth.simulate_read(ref, int(rng.integers(0, len(ref) - L)), L, bool(i & 1), mean, rng, name=f"s{i}",
                            sub_rate=0.01 if i % 3 == 0 else 0.0)
np.savez_compressed("gpurun_out/read1797.npz", dac=r.dac, basecall=np.frombuffer(r.basecall, dtype=np.uint8),
                    refseq=np.frombuffer(r.refseq, dtype=np.uint8), q2r=r.query_to_ref)
p = portbind.Port().normalise(r.raw, r.basecall, r.refseq, r.query_to_ref, mean)
ctx = api.Context(0, keep_debug=True)
ctx.load_model(api.MODEL_PORE, mean)
o = ctx.normaliseEvents([api.Read.from_synth(r, use_dac=True)])[0]
print("DNB_TS_MODE", os.environ.get("DNB_TS_MODE"), "status", o.status, p["status"])
print("rough", o.rough_shift == p["rough_shift"], o.rough_scale == p["rough_scale"])
print("shift", repr(o.shift), repr(p["shift"]), o.shift == p["shift"])
print("scale", repr(o.scale), repr(p["scale"]), o.scale == p["scale"])
print("cleaned equal", np.array_equal(o.cleaned_signal, p["cleaned_signal"]), np.array_equal(o.cleaned_rank, p["cleaned_rank"]), o.cleaned_signal.size)
