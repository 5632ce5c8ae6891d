"""Drop-in test of the C++ shim (dnascent_b200/csrc/shim): the reference's own objects (read constructor, parseCigar,
kmer2index, ...) linked against OUR normaliseEvents / detect_events / probability / llAcrossRead symbols must leave a
DNAscent::read exactly as the unmodified reference does.  Both libraries are prebuilt by oracle/Makefile where the
reference is mounted and travel to the GPU box; nothing here reads /root/reference at run time."""
import numpy as np
import pytest

from dnascent_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def shim(pore_mean):
    from oracle import refbind
    if not refbind.shim_available():
        pytest.skip("oracle/_ref/libdnascent_shim.so not built (needs /root/reference at build time)")
    S = refbind.Ref(shim=True)
    S.set_model(refbind.PORE, pore_mean, np.full(pore_mean.size, 0.14))
    yield S
    S.shutdown()


def _reads(pore_mean, ref, n, seed, lo=1500, hi=9000):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        L = int(rng.integers(lo, hi))
        out.append(synth.simulate_read(ref, int(rng.integers(0, len(ref) - L)), L, bool(i % 2), pore_mean, rng,
                                       name=f"s{i}", sub_rate=0.02 if i % 3 == 0 else 0.0))
    return out


def _same(a: dict, b: dict, tag):
    np.testing.assert_array_equal(a["event_mean"], b["event_mean"], err_msg=tag)
    np.testing.assert_array_equal(a["event_raw_len"], b["event_raw_len"], err_msg=tag)
    np.testing.assert_array_equal(a["align_event"], b["align_event"], err_msg=tag)
    np.testing.assert_array_equal(a["align_kmer"], b["align_kmer"], err_msg=tag)
    for k in ("shift", "scale", "events_per_base", "avg_log_emission", "spanned", "max_gap", "qc_set"):
        assert a[k] == b[k], (tag, k, a[k], b[k])


def test_shim_normalise_events_matches_reference(shim, ref_oracle, pore_mean):
    ref = synth.make_reference(200_000, 31)
    shim.set_reference(ref)
    ref_oracle.set_reference(ref)
    reads = _reads(pore_mean, ref, 12, 32)
    hs = [shim.read_new(r) for r in reads]
    hr = [ref_oracle.read_new(r) for r in reads]
    shim.normalise_batch(hs)                       # one GPU submission for the whole buffer
    n_ok = 0
    for i, (a, b) in enumerate(zip(hs, hr)):
        want = b.normalise(staged=False)
        got = a.outputs(staged=False)
        _same(got, want, f"read {i}")
        # r.events[j].raw concatenated == the reference's
        na = a.L.dnbref_events_raw_concat(a.h, None, 0)
        nb = b.L.dnbref_events_raw_concat(b.h, None, 0)
        assert na == nb
        n_ok += want["align_event"].size > 0
    assert n_ok >= 10
    # the one-read signature (what alignment.cpp:856 / trainCNN.cpp:319 call) gives the same answer
    h1 = shim.read_new(reads[0])
    _same(h1.normalise(staged=False), hr[0].outputs(staged=False), "single-read call")


def test_shim_eventalign_matches_reference(shim, ref_oracle, pore_mean):
    """eventalign(r, 50) through the shim (window chain + builtinViterbi on the GPU, text and r.addSignal on the host)
    against the unmodified reference on the same reads: identical humanReadable_eventalignOut, identical DNN input
    tensors (what r.addSignal left in refCoordToAP, reads.h:288-372)."""
    ref = synth.make_reference(200_000, 41)
    shim.set_reference(ref)
    ref_oracle.set_reference(ref)
    reads = _reads(pore_mean, ref, 8, 42, lo=2000, hi=12000)
    hs = [shim.read_new(r) for r in reads]
    hr = [ref_oracle.read_new(r) for r in reads]
    shim.normalise_batch(hs)
    n = 0
    for i, (a, b) in enumerate(zip(hs, hr)):
        want = b.normalise(staged=False)
        if want["align_event"].size == 0:
            continue
        text_ref = b.eventalign(50)
        text_shim = a.eventalign(50)                 # the one-read signature of alignment.h:22
        assert text_shim == text_ref, f"read {i}"
        pa, pb = a.aligned_positions(), b.aligned_positions()
        for key in ("signal", "core", "residual", "coords"):
            np.testing.assert_array_equal(pa[key], pb[key], err_msg=f"read {i} {key}")
        n += 1
    assert n >= 6


def test_shim_eventalign_features_batch_matches_reference(shim, ref_oracle, pore_mean):
    """Row f2 through the C++ shim: dnb_shim::eventalign_features_batch (window chains + tensors on the device, no text,
    no addSignal) must hand runCNN the vectors the unmodified reference builds from refCoordToAP after its own
    eventalign (reads.h:305-452): signal, core, residual, coordinates, indices, alignment quality -- bit for bit."""
    ref = synth.make_reference(200_000, 51)
    shim.set_reference(ref)
    ref_oracle.set_reference(ref)
    reads = _reads(pore_mean, ref, 8, 52, lo=2000, hi=12000)
    hs = [shim.read_new(r) for r in reads]
    hr = [ref_oracle.read_new(r) for r in reads]
    shim.normalise_batch(hs)
    ok = [i for i, b in enumerate(hr) if b.normalise(staged=False)["align_event"].size > 0]
    assert len(ok) >= 6
    got = shim.eventalign_features_batch([hs[i] for i in ok], 50)
    strands = set()
    for i, g in zip(ok, got):
        hr[i].eventalign(50)
        want = hr[i].aligned_positions()
        assert want["core"].size > 0
        for key in ("signal", "core", "residual", "coords", "ref_index", "query_index", "quality"):
            np.testing.assert_array_equal(g[key], want[key], err_msg=f"read {i} {key}")
        strands.add(bool(hr[i].is_reverse))
    assert strands == {False, True}


def test_shim_resident_chain_matches_reference(shim, ref_oracle, pore_mean):
    """dnb_shim::normalise_eventalign_batch (one device-resident batch: normaliseEvents -> eventalign -> tensors) leaves
    each DNAscent::read as the reference's normaliseEvents does and hands back the reference's DNN inputs; a read that
    fails normalisation comes back with an empty eventAlignment and no tensors (detect.cpp:879-881)."""
    ref = synth.make_reference(200_000, 71)
    shim.set_reference(ref)
    ref_oracle.set_reference(ref)
    reads = _reads(pore_mean, ref, 7, 72, lo=2000, hi=12000)
    rng = np.random.default_rng(73)
    reads.append(synth.simulate_read(ref, 5000, 400, False, pore_mean, rng, name="short"))      # < 1000 cleaned points: QC fail
    hs = [shim.read_new(r) for r in reads]
    hr = [ref_oracle.read_new(r) for r in reads]
    got = shim.eventalign_features_batch(hs, 50, resident=True)
    n_ok = n_fail = 0
    for i, (a, b, g) in enumerate(zip(hs, hr, got)):
        want = b.normalise(staged=False)
        _same(a.outputs(staged=False), want, f"read {i}")
        if want["align_event"].size == 0:
            assert g["core"].size == 0
            n_fail += 1
            continue
        b.eventalign(50)
        ap = b.aligned_positions()
        for key in ("signal", "core", "residual", "coords", "ref_index", "query_index", "quality"):
            np.testing.assert_array_equal(g[key], ap[key], err_msg=f"read {i} {key}")
        n_ok += 1
    assert n_ok >= 5 and n_fail >= 1


def test_shim_detect_events_and_probability(shim, ref_oracle, pore_mean):
    ref = synth.make_reference(50_000, 33)
    r = _reads(pore_mean, ref, 1, 34)[0]
    for x, y in zip(shim.detect_events(r.raw), ref_oracle.detect_events(r.raw)):
        np.testing.assert_array_equal(x, y)
    L, R = shim.L, ref_oracle.L
    for a, b in [(0.5, -3.0), (float("nan"), 1.0), (2.0, float("nan")), (float("nan"), float("nan")), (-700.0, 3.0)]:
        for f in ("dnbref_lnSum", "dnbref_lnProd"):
            x, y = getattr(L, f)(a, b), getattr(R, f)(a, b)
            assert (np.isnan(x) and np.isnan(y)) or x == y
        assert L.dnbref_lnGreaterThan(a, b) == R.dnbref_lnGreaterThan(a, b)
    with pytest.raises(ValueError):
        shim.eln(-1.0)                              # NegativeLog crosses the shim as the reference's exception
    assert np.isnan(shim.eln(0.0)) and shim.eln(2.0) == ref_oracle.eln(2.0)
    assert L.dnbref_normalPDF(0.1, 0.14, 0.3) == R.dnbref_normalPDF(0.1, 0.14, 0.3)


def test_shim_ll_across_read_matches_reference(shim, golden_reads, golden_reference):
    """llAcrossRead through the shim on golden read `read_index` (built by the reference's own read constructor from the
    SAM fields) against the LLRs the unmodified reference produced for it (tests/golden/hmm_v1.npz; the fixture holds
    the unlabelled / BrdU table rows this read touches)."""
    import os
    from oracle import refbind
    h = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hmm_v1.npz"))
    for name_m, name_s, which in (("unl_mean", "unl_stdv", refbind.UNLABELLED), ("ana_mean", "ana_stdv", refbind.ANALOGUE)):
        m, s = np.zeros(4 ** 9), np.zeros(4 ** 9)
        m[h["ranks"]] = h[name_m]
        s[h["ranks"]] = h[name_s]
        shim.set_model(which, m, s)
    shim.shutdown()                                  # the context reloads the three tables on next use
    shim.set_reference(golden_reference)
    g = golden_reads[int(h["read_index"])]
    hs = [shim.read_new(g)]
    shim.normalise_batch(hs)
    np.testing.assert_array_equal(hs[0].outputs(staged=False)["align_event"], g.align[:, 0])
    (pos, llr), = shim.ll_across_read_batch(hs, 12)
    np.testing.assert_array_equal(pos, h["pos_global"])
    np.testing.assert_allclose(llr, h["llr"], rtol=1e-4, atol=1e-6)     # BASELINE.json tolerance for the HMM path
    # the one-read signature llAcrossRead(r, 12) (detect.cpp:883) gives the same calls
    pos1, llr1 = hs[0].ll_across_read(12)
    np.testing.assert_array_equal(pos1, pos)
    np.testing.assert_array_equal(llr1, llr)


def test_shim_resident_hmm_chain_matches_reference(golden_v2, pore_mean):
    """detect --HMM through the C++ shim as one resident chain (dnb_shim::normalise_llAcrossRead_batch): reads built by
    the reference's constructor, normaliseEvents + llAcrossRead on the device, calls against the reference's (1e-4)."""
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
    calls = S.normalise_ll_batch(hs, 12)
    for t, h, (pos, llr) in zip(("a0", "a1"), hs, calls):
        g = reads[t]
        np.testing.assert_array_equal(h.outputs(staged=False)["align_event"], g.align[:, 0], err_msg=t)
        np.testing.assert_array_equal(pos, g.pos_global, err_msg=t)
        np.testing.assert_allclose(llr, g.llr, rtol=1e-4, atol=1e-6, err_msg=t)
    S.shutdown()


def test_shim_one_process_two_gpus(n_cuda, golden_reads, golden_reference, pore_mean):
    """SURVEY s.8(b): one dnb_ctx per process driving several GPUs.  The shim's single context deals whole buffers to
    the least-loaded device; DNAscent::read objects come back identical whichever GPU ran them."""
    from oracle import refbind
    if not refbind.shim_available():
        pytest.skip("oracle/_ref/libdnascent_shim.so not built")
    if n_cuda < 2:
        pytest.skip("needs two CUDA devices")
    import threading
    S = refbind.Ref(shim=True)
    S.set_model(refbind.PORE, pore_mean, np.full(pore_mean.size, 0.14))
    S.shutdown()
    S.set_devices([0, 1])
    S.set_reference(golden_reference)
    before = [S.batches_on_device(k) for k in (0, 1)]
    buffers = [[S.read_new(g) for g in golden_reads] for _ in range(4)]
    threads = [threading.Thread(target=S.normalise_batch, args=(hs,)) for hs in buffers]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for hs in buffers:
        for h, g in zip(hs, golden_reads):
            o = h.outputs(staged=False)
            np.testing.assert_array_equal(o["align_event"], g.align[:, 0])
            np.testing.assert_array_equal(o["align_kmer"], g.align[:, 1])
            assert o["shift"] == g.shift and o["scale"] == g.scale
    used = [S.batches_on_device(k) - b for k, b in zip((0, 1), before)]
    assert sum(used) == 4 and min(used) >= 1, used
    S.shutdown()
    S.set_devices([0])
