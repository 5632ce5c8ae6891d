"""GPU parity tests: the CUDA path, called through the C ABI, against the golden fixtures (outputs of the unmodified
reference) and against the CPU oracle on seeded inputs.  Bit-exact for event boundaries, event means, alignment
paths, QC flags and scalings (integer / IEEE-exact work); 1e-4 relative for the analogue log-likelihoods."""
import numpy as np
import pytest

from dnascent_b200 import api, synth

pytestmark = pytest.mark.gpu


def _check_against_golden(res, g):
    assert res.status == api.READ_OK
    assert res.et_n == g.et_mean.size
    np.testing.assert_array_equal(res.event_mean, g.event_mean)
    np.testing.assert_array_equal(np.diff(res.event_start.astype(np.int64)), g.event_raw_len.astype(np.int64))
    np.testing.assert_array_equal(res.eventAlignment, g.align)
    assert res.rough_shift == g.rough_shift and res.rough_scale == g.rough_scale
    assert res.shift == g.shift and res.scale == g.scale
    assert res.eventsPerBase == g.events_per_base
    assert res.avg_log_emission == g.avg_log_emission
    assert res.spanned == g.spanned and res.maxGap == g.max_gap
    np.testing.assert_array_equal(res.cleaned_signal, g.cleaned_signal)
    np.testing.assert_array_equal(res.cleaned_rank, g.cleaned_rank)


def test_golden_reads_float_input(ctx, golden_reads):
    out = ctx.normaliseEvents([api.Read(g.raw, g.basecall, g.refseq, g.query_to_ref) for g in golden_reads])
    for res, g in zip(out, golden_reads):
        _check_against_golden(res, g)


def test_golden_reads_int16_input(ctx, golden_reads):
    reads = [api.Read(None, g.basecall, g.refseq, g.query_to_ref, dac=g.dac, dac_offset=float(synth.DAC_OFFSET),
                      dac_scale=float(synth.DAC_SCALE)) for g in golden_reads]
    for res, g in zip(ctx.normaliseEvents(reads), golden_reads):
        _check_against_golden(res, g)


def test_detect_events_drop_in(ctx, golden_reads):
    for g in golden_reads[:3]:
        st, ln, mn, sd = ctx.detect_events(g.raw)
        np.testing.assert_array_equal(st, g.et_start.astype(np.uint64))
        np.testing.assert_array_equal(ln, g.et_length)
        np.testing.assert_array_equal(mn, g.et_mean)
        np.testing.assert_array_equal(sd, g.et_stdv)


def _compare_with_port(res, p, tag=""):
    """CUDA result vs the CPU oracle port on the same input: everything bit-exact."""
    assert res.et_n == p["et_n"], tag
    np.testing.assert_array_equal(res.event_mean.astype(np.float64), p["event_mean"], err_msg=tag)
    if p["event_mean"].size:
        np.testing.assert_array_equal(res.event_start, p["event_start"], err_msg=tag)
    if p["status"] == 3:   # undefined in the reference
        assert res.status == api.READ_UNDEFINED, tag
        return
    assert res.status == p["status"], tag
    assert res.rough_shift == p["rough_shift"] and res.rough_scale == p["rough_scale"], tag
    assert res.shift == p["shift"] and res.scale == p["scale"], tag
    assert res.avg_log_emission == p["avg_log_emission"], tag
    assert res.spanned == p["spanned"] and res.maxGap == p["max_gap"], tag
    if p["status"] == 0:
        np.testing.assert_array_equal(res.eventAlignment[:, 0], p["align_event"], err_msg=tag)
        np.testing.assert_array_equal(res.eventAlignment[:, 1], p["align_kmer"], err_msg=tag)
    else:
        assert res.eventAlignment.shape[0] == 0, tag
    if res.cleaned_signal is not None:
        np.testing.assert_array_equal(res.cleaned_signal, p["cleaned_signal"], err_msg=tag)
        np.testing.assert_array_equal(res.cleaned_rank, p["cleaned_rank"], err_msg=tag)


def test_random_reads_vs_oracle(ctx, port, pore_mean):
    """Seeded reads, both strands, with and without substitutions, lengths straddling the 512-sample tiles."""
    ref = synth.make_reference(400_000, 11)
    rng = np.random.default_rng(12)
    lengths = [int(x) for x in rng.integers(1100, 9000, size=20)] + [12000, 15000, 20000, 1024, 2048, 4096]
    reads = []
    for i, L in enumerate(lengths):
        reads.append(synth.simulate_read(ref, int(rng.integers(0, len(ref) - L)), L, bool(i % 2), pore_mean, rng,
                                         name=f"t{i}", sub_rate=0.03 if i % 3 == 0 else 0.0))
    out = ctx.normaliseEvents([api.Read.from_synth(r) for r in reads])
    n_ok = 0
    for i, (r, o) in enumerate(zip(reads, out)):
        p = port.normalise(r.raw, r.basecall, r.refseq, r.query_to_ref, pore_mean)
        _compare_with_port(o, p, tag=f"read {i} L={lengths[i]}")
        n_ok += o.status == api.READ_OK
    assert n_ok >= len(reads) - 2


def test_segmentation_edge_shapes(ctx, port, pore_mean):
    """Signals the simulator never produces: tiny reads, exact tile multiples, noise, constants, a long stall,
    zero / negative levels (exercise the r.events filter, quirk Q3, and the serial fallback)."""
    rng = np.random.default_rng(5)
    ref = synth.make_reference(20_000, 6)
    base = synth.simulate_read(ref, 100, 3000, False, pore_mean, rng)
    sig = base.raw
    cases = []
    for n in (1, 2, 5, 6, 11, 12, 13, 31, 64, 100, 511, 512, 513, 1023, 1024, 1025, 1536, 4096, 4101):
        cases.append(("prefix%d" % n, sig[:n].copy()))
    cases.append(("noise", (90 + 2.0 * rng.standard_normal(5000)).astype(np.float32)))
    cases.append(("const", np.full(3000, 87.25, dtype=np.float32)))
    stall = sig[:6000].copy(); stall[2000:4500] = np.float32(71.5)
    cases.append(("stall", stall))
    neg = sig[:4000].copy() - np.float32(95.0)
    cases.append(("negative_levels", neg))
    zero = sig[:4000].copy(); zero[1000:1400] = 0.0
    cases.append(("zero_stretch", zero))
    steps = np.repeat(rng.normal(90, 12, size=700), 7).astype(np.float32)
    cases.append(("clean_steps", steps))
    reads = [api.Read(raw, base.basecall, base.refseq, base.query_to_ref) for _, raw in cases]
    out = ctx.normaliseEvents(reads)
    for (name, raw), o in zip(cases, out):
        p = port.normalise(raw, base.basecall, base.refseq, base.query_to_ref, pore_mean)
        _compare_with_port(o, p, tag=name)
        st, ln, mn, sd = ctx.detect_events(raw) if p["et_n"] else (None,) * 4
        if p["et_n"]:
            np.testing.assert_array_equal(st, p["et_start"], err_msg=name)
            np.testing.assert_array_equal(mn, p["et_mean"], err_msg=name)
            np.testing.assert_array_equal(sd, p["et_stdv"], err_msg=name)


def _hmm_tables():
    import os
    h = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hmm_v1.npz"))
    tabs = []
    for name in ("unl_mean", "unl_stdv", "ana_mean", "ana_stdv"):
        t = np.zeros(4 ** 9)
        t[h["ranks"]] = h[name]
        tabs.append(t)
    return h, tabs


def test_sequence_probability_golden(ctx):
    """Analogue / thymidine forward probabilities vs the reference's sequenceProbability (detect.cpp:235-378).
    Tolerance 1e-4 relative (BASELINE.json north_star) -- CUDA log/exp instead of glibc."""
    h, (um, us, am, as_) = _hmm_tables()
    ctx.load_model(api.MODEL_UNLABELLED, um, us)
    ctx.load_model(api.MODEL_ANALOGUE, am, as_)
    off = h["sp_obs_off"]
    n = h["sp_out"].shape[0]
    obs = [h["sp_obs"][off[j]:off[j + 1]] for j in range(n)]
    snips = [h["sp_seq"][j].tobytes() for j in range(n)]
    la, lt = ctx.sequence_probability_batch(obs, snips, h["sp_par"][:, 0], h["sp_par"][:, 1], h["sp_par"][:, 2], window=12)
    np.testing.assert_allclose(la, h["sp_out"][:, 0], rtol=1e-4)
    np.testing.assert_allclose(lt, h["sp_out"][:, 1], rtol=1e-4)


def test_ll_across_read_golden(ctx, golden_reads):
    """llAcrossRead (detect.cpp:393-574) on golden read 4: the host gathers each T site's events exactly like the
    reference (port of the gathering is checked on CPU); the device computes both forward passes per site."""
    from oracle import portbind  # the checker: provides the per-site event snippets the reference would build
    h, (um, us, am, as_) = _hmm_tables()
    ctx.load_model(api.MODEL_UNLABELLED, um, us)
    ctx.load_model(api.MODEL_ANALOGUE, am, as_)
    g = golden_reads[int(h["read_index"])]
    res = ctx.normaliseEvents([api.Read(g.raw, g.basecall, g.refseq, g.query_to_ref)])[0]
    sites = api.gather_sites(g.refseq, h["ref_to_query"], bool(h["is_reverse"]), res.eventAlignment,
                             res.event_mean.astype(np.float64), window=12)
    la, lt = ctx.sequence_probability_batch([s[1] for s in sites], [s[2] for s in sites], res.shift, res.scale,
                                            res.eventsPerBase, window=12)
    pos = np.array([s[0] for s in sites], dtype=np.int64)
    glob = (int(h["ref_end"]) - pos - 1) if bool(h["is_reverse"]) else int(h["ref_start"]) + pos
    order = np.argsort(glob)
    np.testing.assert_array_equal(glob[order], h["pos_global"])
    np.testing.assert_allclose((la - lt)[order], h["llr"], rtol=1e-4, atol=1e-6)
    del portbind


def test_large_batch_vs_oracle_statistics(ctx, port, pore_mean):
    """A few hundred reads in one batch (exercises ordering, tiling, workspace offsets); a random subset is checked
    bit-for-bit against the oracle, every read must come back with a definite status."""
    ref = synth.make_reference(600_000, 21)
    rng = np.random.default_rng(22)
    lengths = synth.lognormal_lengths(300, 6000, rng, lo=1100, hi=40_000)
    reads = synth.simulate_batch(ref, lengths, pore_mean, seed=23, sub_rate=0.01)
    out = ctx.normaliseEvents([api.Read.from_synth(r, use_dac=True) for r in reads])
    assert all(o.status in (0, 1, 2, 3) for o in out)
    assert sum(o.status == 0 for o in out) > 0.9 * len(out)
    for i in rng.choice(len(reads), size=40, replace=False):
        r = reads[i]
        p = port.normalise(r.raw, r.basecall, r.refseq, r.query_to_ref, pore_mean)
        _compare_with_port(out[i], p, tag=f"read {i}")


def _check_v2(res, g, tag):
    assert res.status == api.READ_OK, tag
    assert res.et_n == g.et_n, tag
    np.testing.assert_array_equal(res.event_mean, g.event_mean, err_msg=tag)
    np.testing.assert_array_equal(np.diff(res.event_start.astype(np.int64)), g.event_raw_len.astype(np.int64), err_msg=tag)
    np.testing.assert_array_equal(res.eventAlignment, g.align, err_msg=tag)
    assert (res.rough_shift, res.rough_scale, res.shift, res.scale) == (g.rough_shift, g.rough_scale, g.shift, g.scale), tag
    assert res.eventsPerBase == g.events_per_base and res.avg_log_emission == g.avg_log_emission, tag
    assert res.spanned == g.spanned and res.maxGap == g.max_gap, tag
    np.testing.assert_array_equal(res.cleaned_signal, g.cleaned_signal, err_msg=tag)
    np.testing.assert_array_equal(res.cleaned_rank, g.cleaned_rank, err_msg=tag)


def test_golden_v2_indel_cigars_and_analogue_reads(ctx, golden_v2):
    """reads_v2.npz: queryToRef from CIGARs with insertions / deletions / soft clips (duplicate and out-of-range
    entries, as parseCigar produces them) and BrdU-substituted signal; bit-exact against the reference's outputs."""
    reads, _, _ = golden_v2
    tags = list(reads)
    out = ctx.normaliseEvents([api.Read(reads[t].raw, reads[t].basecall, reads[t].refseq, reads[t].query_to_ref) for t in tags])
    for t, res in zip(tags, out):
        _check_v2(res, reads[t], t)


def test_analogue_llr_configs3(golden_v2, pore_mean):
    """BASELINE.json configs[3]: analogue log-likelihood ratios of BrdU-substituted reads, through the C++ shim
    (reference read constructor + our normaliseEvents + our llAcrossRead) against the reference's LLRs, 1e-4 relative."""
    from oracle import refbind
    if not refbind.shim_available():
        pytest.skip("oracle/_ref/libdnascent_shim.so not built")
    reads, (um, us, am, as_), reference = golden_v2
    S = refbind.Ref(shim=True)
    S.set_model(refbind.PORE, pore_mean, np.full(pore_mean.size, 0.14))
    S.set_model(refbind.UNLABELLED, um, us)
    S.set_model(refbind.ANALOGUE, am, as_)
    S.shutdown()
    S.set_reference(reference)
    hs = [S.read_new(reads[t]) for t in ("a0", "a1")]
    S.normalise_batch(hs)
    calls = S.ll_across_read_batch(hs, 12)
    for t, h, (pos, llr) in zip(("a0", "a1"), hs, calls):
        g = reads[t]
        np.testing.assert_array_equal(h.outputs(staged=False)["align_event"], g.align[:, 0], err_msg=t)
        np.testing.assert_array_equal(pos, g.pos_global, err_msg=t)
        np.testing.assert_allclose(llr, g.llr, rtol=1e-4, atol=1e-6, err_msg=t)
    S.shutdown()


def test_ultra_long_read_configs2(ctx, port, pore_mean):
    """BASELINE.json configs[2] (100 kb - 1 Mb reads): long band walks, >1000 segmentation tiles per read, trace
    windows far apart.  One 130-kb and one 260-kb read against the CPU oracle, bit-exact."""
    ref = synth.make_reference(400_000, 41)
    rng = np.random.default_rng(42)
    reads = [synth.simulate_read(ref, 1000, 130_000, False, pore_mean, rng, name="u0"),
             synth.simulate_read(ref, 90_000, 260_000, True, pore_mean, rng, name="u1", sub_rate=0.01)]
    out = ctx.normaliseEvents([api.Read.from_synth(r, use_dac=True) for r in reads])
    for r, o in zip(reads, out):
        p = port.normalise(r.raw, r.basecall, r.refseq, r.query_to_ref, pore_mean)
        _compare_with_port(o, p, tag=r.name)
        assert o.status == api.READ_OK and o.eventAlignment.shape[0] > len(r.basecall)


def test_one_megabase_read_configs2(ctx_compact, port, pore_mean):
    """BASELINE.json configs[2], upper end: ONE 1-Mb read (1.25*10^7 samples, ~2.4*10^4 segmentation tiles, ~3.2*10^6
    bands = a 100-MB trace walked back in 5*10^4 rounds) against the CPU oracle, bit-exact, through the compact wire
    format (event lengths and 2-bit steps of a 3*10^6-step path)."""
    ref = synth.make_reference(1_050_000, 51)
    rng = np.random.default_rng(52)
    r = synth.simulate_read(ref, 20_000, 1_000_000, True, pore_mean, rng, name="mb1")
    assert r.dac.size > 1.0e7
    o = ctx_compact.normaliseEvents([api.Read.from_synth(r, use_dac=True).with_runs()])[0]
    p = port.normalise(r.raw, r.basecall, r.refseq, r.query_to_ref, pore_mean)
    _compare_with_port(o, p, tag="mb1")
    assert o.status == api.READ_OK and o.eventAlignment.shape[0] > 2_000_000


def test_int16_ingest_with_dorado_slice(ctx, port, pore_mean):
    """Row f3: a POD5 record longer than the read (Dorado-trimmed prefix, a split sibling's samples after it) goes in
    as int16 through dnb_dorado_slice's pointer offset; results must equal the oracle on the reference's own
    erase-then-process order (pod5.cpp:56-93 then normaliseEvents)."""
    ref = synth.make_reference(60_000, 61)
    reads = synth.simulate_batch(ref, [3000, 4500], pore_mean, seed=62)
    rng = np.random.default_rng(63)
    ins, want = [], []
    for j, sr in enumerate(reads):
        pre, post = int(rng.integers(50, 4000)), int(rng.integers(0, 3000))
        junk = lambda n: rng.integers(200, 900, n).astype(np.int16)
        record = np.concatenate([junk(pre), sr.dac, junk(post)])
        if j == 0:      # unsplit read: ts = trimmed prefix, ns = end of the read's samples
            sl = api.dorado_slice(record.size, signal_length=pre + sr.dac.size, signal_trim=pre)
        else:           # split read: sp = start inside the parent record, ts on top of it
            sl = api.dorado_slice(record.size, signal_length=40 + sr.dac.size, signal_trim=40, signal_start_coord=pre - 40,
                                  is_split=True)
        assert (sl.start, sl.stop) == (pre, pre + sr.dac.size)
        ins.append(api.Read(None, sr.basecall, sr.refseq, sr.query_to_ref, dac=record[sl],
                            dac_offset=float(synth.DAC_OFFSET), dac_scale=float(synth.DAC_SCALE)))
        raw = synth.dac_to_pa(record).astype(np.float64)[sl]             # convert everything, then erase (the reference's order)
        want.append(port.normalise(raw, sr.basecall, sr.refseq, sr.query_to_ref, pore_mean))
    for o, p in zip(ctx.normaliseEvents(ins), want):
        assert o.status == p["status"] == 0
        np.testing.assert_array_equal(o.event_mean.astype(np.float64), p["event_mean"])
        np.testing.assert_array_equal(o.eventAlignment[:, 0], p["align_event"])
        np.testing.assert_array_equal(o.eventAlignment[:, 1], p["align_kmer"])
        assert o.shift == p["shift"] and o.scale == p["scale"]


@pytest.mark.parametrize("fixture", ["read_theilsen_nan_slope.npz", "read_theilsen_nan_slope_b.npz"])
def test_theil_sen_with_a_nan_slope(ctx, pore_mean, fixture):
    """Two cleaned points with identical signal and identical model level give a 0/0 slope, and std::sort with a NaN in
    the range is outside its contract: which slope the reference takes as the median depends on where libstdc++'s
    introsort leaves the NaN -- behind the median on the first read (found by scripts/ea_statistical_parity.py), in
    front of it on the second.  The device follows the NaN through the introsort (theilsen.cu, nan_sort_path.cuh).  The
    fixtures are those reads with the unmodified reference's scalings (tests/golden/make_golden_nan_slope.py)."""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fixture))
    r = api.Read(None, d["basecall"].tobytes(), d["refseq"].tobytes(), d["q2r"], dac=d["dac"],
                 dac_offset=float(synth.DAC_OFFSET), dac_scale=float(synth.DAC_SCALE))
    o = ctx.normaliseEvents([r])[0]
    assert o.status == api.READ_OK and o.eventAlignment.shape[0] == int(d["n_align"])
    assert o.rough_shift == float(d["rough_shift"]) and o.rough_scale == float(d["rough_scale"])
    assert o.shift == float(d["shift"]) and o.scale == float(d["scale"])


def test_theil_sen_batch_with_nan_slopes(ctx, port, pore_mean):
    """dnb_theil_sen_batch (estimateScaling_theilSen alone) on 48 synthetic inputs with none, one or two 0/0 slopes
    (tests/helpers/nan_slope_cases.py) against the port's literal std::sort, which tests/test_oracle_golden.py pins to
    the unmodified reference on the same cases.  With ONE NaN the device must give the reference's answer whichever
    side of the median the introsort leaves it; two NaNs (about 10^-6 of real reads) are not emulated: there the
    device's answer is the "NaN sorted last" one, checked as such."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
    from nan_slope_cases import cases
    cs = cases(pore_mean, 48)
    shift, scale = ctx.theil_sen_batch([c[0] for c in cs], [c[1] for c in cs], [c[2] for c in cs], [c[3] for c in cs])
    moved = 0
    for k, (sig, rk, sh, sc) in enumerate(cs):
        want = port.theil_sen(sig, rk, pore_mean, sh, sc)
        n = sig.size
        skip = (n - 100) // 1000 if n - 100 > 1000 else 1
        idx = 50 + np.arange(min(n - 100, 1000)) * skip
        x, y = (sig[idx] - sh) / sc, pore_mean[rk[idx]]
        iu = np.triu_indices(idx.size, 1)
        with np.errstate(all="ignore"):
            sl = (y[iu[0]] - y[iu[1]]) / (x[iu[0]] - x[iu[1]])
        n_nan = int(np.isnan(sl).sum())
        if n_nan <= 1:
            assert (shift[k], scale[k]) == want, (k, n_nan)
            if n_nan == 1:
                nan_last = sc * (1.0 / np.sort(sl[~np.isnan(sl)])[sl.size // 2])
                moved += scale[k] != nan_last
        else:
            assert scale[k] == sc * (1.0 / np.sort(sl[~np.isnan(sl)])[sl.size // 2]), (k, n_nan)
    assert moved >= 10      # the NaN really is in front of the median in a good share of the cases
