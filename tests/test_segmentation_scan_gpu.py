"""The streaming warp scan of the (sum, sumsq) checkpoints (csrc/seg_scan.cu) on signals built to stress it: sums that
cross many binades, addends with few mantissa bits (exact rounding ties), spikes, zeros and negative levels, a read
whose `sum` cannot be exact in double (must fall back to the literal serial kernel).  Event boundaries and means are
compared bit for bit with the CPU oracle; the same signals go through the one-lane checkpoint chain in a second
process-wide mode only via DNB_SEG_PARITY_SCAN=0 (the acceptance run of scripts/)."""
import numpy as np
import pytest

from dnascent_b200 import api, synth
from test_gpu_parity import _compare_with_port

pytestmark = pytest.mark.gpu


def _signals(rng):
    def pod5_like(n, scale=0.1755, off=-240.0):
        lvl = np.repeat(rng.normal(650, 60, size=n // 8 + 1), 8)[:n]
        dac = np.rint(lvl + rng.normal(0, 6, size=n)).astype(np.int16)
        return ((dac.astype(np.float32) + np.float32(off)) * np.float32(scale)).astype(np.float32)

    steps = np.repeat(rng.normal(90, 12, size=6000), 7).astype(np.float32)
    return {
        "pod5_like": pod5_like(70_000),
        "ties": (np.round((steps + rng.normal(0, 1.5, size=steps.size)) * 4) / 4).astype(np.float32),
        "powers_of_two": np.ldexp(1.0, rng.integers(3, 9, size=30_000)).astype(np.float32),
        "growing": (np.linspace(1, 3000, 50_000) * rng.uniform(0.9, 1.1, size=50_000)).astype(np.float32),
        "spikes": np.where(rng.random(40_000) < 0.01, 30000.0, np.repeat(rng.normal(90, 10, size=5000), 8)).astype(np.float32),
        "zeros_and_negatives": np.where(rng.random(30_000) < 0.2, 0.0, np.repeat(rng.normal(0, 50, size=3750), 8)).astype(np.float32),
        "tiny_then_large": np.concatenate([np.full(3000, 2.0 ** -19, dtype=np.float32), pod5_like(20_000)]),
        "inexact_sum": np.concatenate([np.full(4000, 2.0 ** -20 * 1.5, dtype=np.float32),
                                       np.full(4000, 2.0 ** 29 * 1.25, dtype=np.float32), pod5_like(4000)]),
        "exact_2048": pod5_like(2048), "one_short_of_a_chunk": pod5_like(2047), "one_past_a_chunk": pod5_like(2049),
    }


def test_adversarial_signals_segment_like_the_oracle(ctx, port, pore_mean):
    rng = np.random.default_rng(77)
    ref = synth.make_reference(20_000, 6)
    base = synth.simulate_read(ref, 100, 3000, False, pore_mean, rng)
    sig = _signals(rng)
    out = ctx.normaliseEvents([api.Read(x, base.basecall, base.refseq, base.query_to_ref) for x in sig.values()])
    for (name, x), o in zip(sig.items(), out):
        p = port.normalise(x, base.basecall, base.refseq, base.query_to_ref, pore_mean)
        _compare_with_port(o, p, tag=name)


def test_long_read_checkpoints(ctx, port, pore_mean):
    """A 400-kb read (5*10^6 samples, ~2400 chunks, ~45 binade crossings of the sumsq chain) through the scan."""
    ref = synth.make_reference(450_000, 43)
    rng = np.random.default_rng(44)
    r = synth.simulate_read(ref, 2000, 400_000, False, pore_mean, rng, name="u400")
    o = ctx.normaliseEvents([api.Read.from_synth(r, use_dac=True)])[0]
    p = port.normalise(r.raw, r.basecall, r.refseq, r.query_to_ref, pore_mean)
    _compare_with_port(o, p, tag="u400")
    assert o.status == api.READ_OK
