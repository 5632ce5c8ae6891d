"""The compact wire formats either side of the pipeline (csrc/pack.cu) against the dense ones and the goldens:
queryToRef as runs (host -> device), events as u8 lengths + escapes and the alignment as a 2-bit path (device ->
host), the signal DMA'd straight out of page-locked caller memory.  A re-encoding must not change a bit."""
import numpy as np
import pytest

from dnascent_b200 import api, synth, _lib
from test_gpu_parity import _check_against_golden, _compare_with_port

pytestmark = pytest.mark.gpu


def _dac_read(g, runs=False):
    r = api.Read(None, g.basecall, g.refseq, g.query_to_ref, dac=g.dac, dac_offset=float(synth.DAC_OFFSET),
                 dac_scale=float(synth.DAC_SCALE))
    return r.with_runs() if runs else r


def test_compact_results_match_goldens(ctx_compact, golden_reads):
    for res, g in zip(ctx_compact.normaliseEvents([_dac_read(g) for g in golden_reads]), golden_reads):
        _check_against_golden(res, g)


def test_q2r_runs_match_goldens(ctx, ctx_compact, golden_reads, golden_v2):
    """`{L}M` reads (one run each) and the indel / soft-clip CIGARs of reads_v2 (many runs, stride 0 and 1)."""
    reads_v2 = golden_v2[0]
    gs = list(golden_reads) + [reads_v2["i0"], reads_v2["i1"]]
    for c in (ctx, ctx_compact):
        dense = c.normaliseEvents([_dac_read(g) for g in gs])
        runs = c.normaliseEvents([_dac_read(g, runs=True) for g in gs])
        for a, b, g in zip(dense, runs, gs):
            assert a.status == b.status == api.READ_OK, g.name
            np.testing.assert_array_equal(a.eventAlignment, b.eventAlignment)
            np.testing.assert_array_equal(a.eventAlignment, g.align)
            np.testing.assert_array_equal(a.cleaned_rank, b.cleaned_rank)
            np.testing.assert_array_equal(b.cleaned_rank, g.cleaned_rank)
            np.testing.assert_array_equal(b.cleaned_signal, g.cleaned_signal)
            assert a.shift == b.shift == g.shift and a.scale == b.scale == g.scale


def test_runs_of_random_cigars():
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(1, 400))
        q2r = np.full(n, -1, dtype=np.int32)
        pos, i = int(rng.integers(0, 50)), 0
        while i < n:
            kind, ln = int(rng.integers(0, 4)), int(rng.integers(1, 30))
            ln = min(ln, n - i)
            if kind == 0:
                q2r[i:i + ln] = pos + np.arange(ln); pos += ln
            elif kind == 1:
                q2r[i:i + ln] = pos
            elif kind == 2:
                pos += ln; ln = 0
            i += ln
        runs = api.q2r_to_runs(q2r)
        back = np.full(n, -1, dtype=np.int64)
        for a, ln, rs, st in runs:
            back[a:a + ln] = rs + st * np.arange(ln)
        np.testing.assert_array_equal(back, q2r)


def test_pinned_caller_memory_is_dma_source(ctx_compact, golden_reads):
    """All reads of one submission live in one page-locked buffer (dnb_host_register): the library copies them to the
    device from there (no staging pass) and the result is the golden one."""
    total = sum(g.dac.size for g in golden_reads) + 7 * len(golden_reads)
    buf = np.zeros(total, dtype=np.int16)
    api.host_register(buf)
    try:
        reads, at = [], 0
        for k, g in enumerate(golden_reads):
            at += k % 7                                   # odd offsets: sources need no alignment
            view = buf[at:at + g.dac.size]
            view[:] = g.dac
            at += g.dac.size
            reads.append(api.Read(None, g.basecall, g.refseq, g.query_to_ref, dac=view,
                                  dac_offset=float(synth.DAC_OFFSET), dac_scale=float(synth.DAC_SCALE)).with_runs())
        b = ctx_compact.submit(reads)
        h2d, d2h = b.io_bytes()
        out = b.results()
        b.release()
        for res, g in zip(out, golden_reads):
            _check_against_golden(res, g)
        n_samples = sum(g.dac.size for g in golden_reads)
        n_bases = sum(len(g.basecall) for g in golden_reads)
        # 2 B per sample + the sequences + small per-read tables; no dense queryToRef (4 B per base)
        assert h2d < 2 * n_samples + 2.2 * n_bases + 4096 * len(golden_reads)
        n_events = sum(g.event_mean.size for g in golden_reads)
        assert d2h < 5 * n_events + n_events // 2 + 4096 * len(golden_reads)     # dense would be 16 B per event
    finally:
        api.host_unregister(buf)


def test_compact_edge_shapes_and_escapes(ctx_compact, port, pore_mean):
    """Events longer than 254 samples (a stall) take the escape path of the u8 length coding; failed and undefined reads
    come back empty in both formats."""
    rng = np.random.default_rng(5)
    ref = synth.make_reference(20_000, 6)
    base = synth.simulate_read(ref, 100, 3000, False, pore_mean, rng)
    sig = base.raw
    cases = [("prefix%d" % n, sig[:n].copy()) for n in (1, 6, 13, 512, 1025, 4101)]
    stall = sig[:9000].copy(); stall[2000:4500] = np.float32(71.5); stall[6000:6300] = np.float32(80.25)
    cases.append(("stall", stall))
    cases.append(("const", np.full(3000, 87.25, dtype=np.float32)))
    cases.append(("whole", sig.copy()))
    out = ctx_compact.normaliseEvents([api.Read(raw, base.basecall, base.refseq, base.query_to_ref) for _, raw in cases])
    long_events = 0
    for (name, raw), o in zip(cases, out):
        p = port.normalise(raw, base.basecall, base.refseq, base.query_to_ref, pore_mean)
        _compare_with_port(o, p, tag=name)
        if o.event_start.size:
            long_events += int((np.diff(o.event_start.astype(np.int64)) >= 255).sum())
    assert long_events >= 2


def test_random_reads_compact_vs_oracle(ctx_compact, port, pore_mean):
    ref = synth.make_reference(300_000, 21)
    rng = np.random.default_rng(22)
    lengths = [int(x) for x in rng.integers(1100, 9000, size=10)] + [12000, 1024]
    reads = [synth.simulate_read(ref, int(rng.integers(0, len(ref) - L)), L, bool(i % 2), pore_mean, rng, name=f"c{i}",
                                 sub_rate=0.03 if i % 3 == 0 else 0.0) for i, L in enumerate(lengths)]
    out = ctx_compact.normaliseEvents([api.Read.from_synth(r, use_dac=True).with_runs() for r in reads])
    for i, (r, o) in enumerate(zip(reads, out)):
        _compare_with_port(o, port.normalise(r.raw, r.basecall, r.refseq, r.query_to_ref, pore_mean), tag=f"read {i}")


def test_resident_chain_on_compact_format(ctx_compact, golden_reads):
    """dnb_submit_chain with the compact result format: eventalign still reads dense pairs, built in HBM only."""
    from oracle import portbind
    P = portbind.Port()
    mean = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden",
                                              "pore_model_r10.4.1_400bps.npz"))["mean"].astype(np.float64)
    gs = golden_reads[:3]
    extra = []
    for i, g in enumerate(gs):
        q2r = np.asarray(g.query_to_ref)
        r2q = np.zeros(len(g.refseq), dtype=np.int32)
        r2q[q2r[q2r >= 0]] = np.nonzero(q2r >= 0)[0]
        extra.append(dict(ref_to_query=r2q, is_reverse=bool(i & 1), ref_start=0, ref_end=len(g.refseq)))
    b = ctx_compact.submit_chain([_dac_read(g, runs=True) for g in gs], extra, 50, want_records=True)
    res, feats = b.results(), b.feature_results(want_records=True)
    b.release()
    for g, o, x, f in zip(gs, res, extra, feats):
        _check_against_golden(o, g)
        rec = P.eventalign(g.refseq, x["ref_to_query"], g.align[:, 0], g.align[:, 1], g.event_mean.astype(np.float64),
                           g.shift, g.scale, g.events_per_base, mean)
        assert f["status"] == 0
        for key in ("event", "ref_pos", "label", "indel"):
            np.testing.assert_array_equal(f[key], rec[key], err_msg=key)


def test_one_context_two_devices(n_cuda, pore_mean, golden_reads):
    """dnb_config.devices: one context (one process) deals submissions to the least-loaded GPU; results are the goldens
    whichever device ran them."""
    if n_cuda < 2:
        pytest.skip("needs two CUDA devices")
    c = api.Context(devices=[0, 1], keep_debug=True, result_format=api.RESULT_COMPACT)
    c.load_model(api.MODEL_PORE, pore_mean)
    try:
        batches = [c.submit([_dac_read(g, runs=True) for g in golden_reads]) for _ in range(4)]
        used = {b.device() for b in batches}
        for b in batches:
            for res, g in zip(b.results(), golden_reads):
                _check_against_golden(res, g)
        assert used == {0, 1}
        for b in batches:
            b.release()
    finally:
        c.close()
