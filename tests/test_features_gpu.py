"""SURVEY s.8 row f2 on the GPU: dnb_eventalign_features_batch (dnascent_b200/csrc/features.cu) through the C ABI.
The DNN input tensors -- signal [P,20], core / residual k-mer indices, reference coordinates / indices, query indices,
alignment quality -- must be BIT-IDENTICAL to what the unmodified reference's makeSignalTensor & co. (reads.h:305-452)
returned for the golden reads (tests/golden/eventalign_v1.npz, keys e_<tag>_ap_*), and to the CPU oracle port on fresh
seeded reads (forward and reverse strand, int16 DAC and float32 pA input, with and without called positions)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from test_eventalign_cpu import AP_KEYS, all_golden_reads, event_starts, golden_eventalign_inputs, golden_records
from dnascent_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ea_golden():
    return np.load(os.path.join(GOLDEN, "eventalign_v1.npz"))


def golden_feature_inputs(g, e, tag, **extra):
    p = f"e_{tag}_"
    ref_start, ref_end, is_rev = (int(x) for x in e[p + "strand"])
    d = golden_eventalign_inputs(g, e, tag)
    d.update(raw=g.raw.astype(np.float32), event_start=event_starts(g.event_raw_len), is_reverse=is_rev,
             ref_start=ref_start, ref_end=ref_end)
    d.update(extra)
    return d


def test_features_golden_tensors(ctx, ea_golden, golden_reads, golden_v2):
    e = ea_golden
    reads = all_golden_reads(golden_reads, golden_v2)
    out = ctx.eventalign_features([golden_feature_inputs(g, e, tag) for tag, g in reads], window=50)
    n_rev = 0
    for (tag, g), o in zip(reads, out):
        p = f"e_{tag}_"
        assert o["status"] == api.READ_OK, tag
        for key in ("event", "ref_pos", "label", "indel"):
            np.testing.assert_array_equal(o[key], e[p + key], err_msg=f"{tag} {key}")
        for key in AP_KEYS:
            np.testing.assert_array_equal(o[key], e[p + "ap_" + key], err_msg=f"{tag} {key}")
        assert o["signal"].dtype == np.float32 and o["core"].dtype == np.float32 and o["residual"].dtype == np.float32
        assert o["signal"].shape[1] == 20 and o["signal"].shape[0] > 0.8 * (len(g.refseq) - 8)
        n_rev += int(e[p + "strand"][2])
    assert 0 < n_rev < len(reads)          # both makeSignalTensor iteration orders are covered


def test_features_called_positions_and_no_records(ctx, ea_golden, golden_reads, golden_v2):
    """alignment.cpp:711: coordinates that already have a call are not added; records may stay on the device."""
    e = ea_golden
    reads = all_golden_reads(golden_reads, golden_v2)
    called = [np.sort(e[f"e_{tag}_ap_coords"][::3]) for tag, _ in reads]
    out = ctx.eventalign_features([golden_feature_inputs(g, e, tag, called=c) for (tag, g), c in zip(reads, called)],
                                  window=50, want_records=False)
    for (tag, g), c, o in zip(reads, called, out):
        p = f"e_{tag}_"
        assert o["status"] == api.READ_OK and "event" not in o
        keep = ~np.isin(e[p + "ap_coords"], c)
        for key in AP_KEYS:
            np.testing.assert_array_equal(o[key], e[p + "ap_" + key][keep], err_msg=f"{tag} {key}")


def test_features_after_normalise_vs_port(ctx, port, pore_mean):
    """The read loop's order (detect.cpp:876-888 then runCNN's tensors): product normaliseEvents -> product
    eventalign + features, against the port fed the product's records; int16 DAC and float32 input, both strands."""
    ref = synth.make_reference(200_000, 191)
    base = synth.simulate_batch(ref, [9000, 12000, 700, 15000, 10000, 8000, 30000, 2000, 45000], pore_mean, seed=192,
                                sub_rate=0.01)
    res = ctx.normaliseEvents([api.Read.from_synth(r, use_dac=True) for r in base])
    reads, strands = [], []
    for i, (sr, o) in enumerate(zip(base, res)):
        if o.status != api.READ_OK:
            continue
        r2q = np.zeros(len(sr.refseq), dtype=np.int32)
        q2r = np.asarray(sr.query_to_ref)
        r2q[q2r[q2r >= 0]] = np.nonzero(q2r >= 0)[0]
        is_rev = bool(i & 1)
        d = dict(refseq=sr.refseq, ref_to_query=r2q, eventAlignment=o.eventAlignment, event_mean=o.event_mean,
                 shift=o.shift, scale=o.scale, events_per_base=o.eventsPerBase, event_start=o.event_start,
                 is_reverse=is_rev, ref_start=1000 * i, ref_end=1000 * i + len(sr.refseq))
        if i % 3 == 0:
            d["raw"] = sr.raw.astype(np.float32)
        else:
            d.update(raw_dac=sr.dac, dac_offset=float(synth.DAC_OFFSET), dac_scale=float(synth.DAC_SCALE))
        reads.append((sr, d))
        strands.append(is_rev)
    assert len(reads) >= 6 and any(strands) and not all(strands)
    out = ctx.eventalign_features([d for _, d in reads], window=50)
    assert ctx.features_last_kernel_ms() > 0.0
    for (sr, d), o in zip(reads, out):
        assert o["status"] == api.READ_OK
        rec = {k: o[k] for k in ("event", "ref_pos", "label", "indel")}
        exp = port.dnn_features(d["refseq"], d["ref_to_query"], d["is_reverse"], d["ref_start"], d["ref_end"], rec,
                                sr.raw.astype(np.float64), d["event_start"], d["shift"], d["scale"])
        assert exp["signal"].shape[0] > 0.8 * (len(sr.refseq) - 8)
        for key in AP_KEYS:
            np.testing.assert_array_equal(o[key], exp[key], err_msg=key)
        # size-independent properties: rows ascend along the read, one row per reference position, zero padding is a suffix
        assert np.all(np.diff(o["ref_index"].astype(np.int64)) > 0)
        step = np.diff(o["coords"].astype(np.int64))
        assert np.all(step < 0) if d["is_reverse"] else np.all(step > 0)
        nz = o["signal"] != 0
        assert np.all(nz[:, :-1] | ~nz[:, 1:])
        assert 1 <= o["core"].min() and o["core"].max() <= 1024 and 1 <= o["residual"].min() and o["residual"].max() <= 256


def test_features_edge_cases(ctx, pore_mean):
    seq = synth.make_reference(200, 5)
    base = dict(refseq=seq, ref_to_query=np.arange(200, dtype=np.int32), eventAlignment=np.zeros((0, 2), dtype=np.uint32),
                event_mean=np.zeros(0, dtype=np.float32), shift=90.0, scale=15.0, events_per_base=2.0,
                raw=np.zeros(16, dtype=np.float32), event_start=np.zeros(1, dtype=np.uint32), is_reverse=False,
                ref_start=0, ref_end=200)
    undefined = dict(base, events_per_base=1.0)
    al = np.stack([np.arange(300), np.arange(300) * 192 // 300], axis=1).astype(np.uint32)
    live = dict(base, eventAlignment=al, event_mean=np.full(300, 95.0, dtype=np.float32),
                raw=np.full(1500, 95.0, dtype=np.float32), event_start=(np.arange(301) * 5).astype(np.uint32))
    tight = dict(live, row_capacity=3)                                    # too few rows -> OVERFLOW, nothing written past them
    bad_es = dict(live, event_start=(np.arange(301) * 6).astype(np.uint32))   # events address samples past the signal
    out = ctx.eventalign_features([base, undefined, live, tight, bad_es], window=50)
    assert out[0]["status"] == api.READ_OK and out[0]["signal"].shape == (0, 20)
    assert out[1]["status"] == api.READ_UNDEFINED and out[1]["signal"].shape[0] == 0
    assert out[2]["status"] == api.READ_OK and out[2]["signal"].shape[0] > 0
    # a flat 95 pA signal scales to (float)((95-90)/15) everywhere a sample exists
    v = np.float32((95.0 - 90.0) / 15.0)
    assert np.all((out[2]["signal"] == v) | (out[2]["signal"] == 0))
    assert out[3]["status"] == api.READ_OVERFLOW and out[3]["signal"].shape[0] == 0
    assert out[4]["status"] == api.READ_UNDEFINED
    assert ctx.eventalign_features([], window=50) == []


def test_resident_chain_golden_tensors(ctx, ea_golden, golden_reads, golden_v2):
    """The whole read loop on the device (detect.cpp:876-888 + runCNN's inputs): int16 DAC in, DNN tensors out, nothing
    in between crosses PCIe but refToQuery.  normaliseEvents, eventalign and the tensor builder chained on the batch's
    resident arrays must reproduce the unmodified reference's tensors and records for every golden read."""
    e = ea_golden
    reads = all_golden_reads(golden_reads, golden_v2)
    b = ctx.upload([api.Read(None, g.basecall, g.refseq, g.query_to_ref, dac=g.dac, dac_offset=float(synth.DAC_OFFSET),
                             dac_scale=float(synth.DAC_SCALE)) for _, g in reads])
    b.run()
    extra = []
    for tag, g in reads:
        ref_start, ref_end, is_rev = (int(x) for x in e[f"e_{tag}_strand"])
        extra.append(dict(ref_to_query=e[f"e_{tag}_ref_to_query"], is_reverse=is_rev, ref_start=ref_start, ref_end=ref_end))
    out = b.eventalign_features(extra, window=50, want_records=True)
    t = b.stage2_timings()
    assert t["eventalign_kernel_ms"] > 0 and t["features_kernel_ms"] > 0 and t["d2h_bytes"] > 0
    for (tag, g), o in zip(reads, out):
        p = f"e_{tag}_"
        assert o["status"] == api.READ_OK, tag
        for key in ("event", "ref_pos", "label", "indel"):
            np.testing.assert_array_equal(o[key], e[p + key], err_msg=f"{tag} {key}")
        for key in AP_KEYS:
            np.testing.assert_array_equal(o[key], e[p + "ap_" + key], err_msg=f"{tag} {key}")
    # the normaliseEvents results of the same batch are still there, and a second call (float input would be a new batch)
    res = b.results()
    for (tag, g), r in zip(reads, res):
        np.testing.assert_array_equal(r.eventAlignment, g.align, err_msg=tag)
    out2 = b.eventalign_features(extra, window=50)                  # idempotent, records left on the device
    for o, o2 in zip(out, out2):
        assert "event" not in o2
        for key in AP_KEYS:
            np.testing.assert_array_equal(o[key], o2[key])
    b.release()


def test_resident_chain_matches_host_array_form(ctx, pore_mean):
    """Float32 input, failed reads in the batch, called positions: resident form == dnb_eventalign_features_batch."""
    ref = synth.make_reference(150_000, 291)
    base = synth.simulate_batch(ref, [6000, 300, 9000, 12000, 2500, 20000], pore_mean, seed=292, sub_rate=0.01)
    b = ctx.upload([api.Read.from_synth(r) for r in base])
    b.run()
    b.fetch()
    res = b.results()
    assert any(o.status != api.READ_OK for o in res) and sum(o.status == api.READ_OK for o in res) >= 4
    extra, host_in = [], []
    for i, (sr, o) in enumerate(zip(base, res)):
        r2q = np.zeros(len(sr.refseq), dtype=np.int32)
        q2r = np.asarray(sr.query_to_ref)
        r2q[q2r[q2r >= 0]] = np.nonzero(q2r >= 0)[0]
        called = np.arange(500 * i + 40, 500 * i + len(sr.refseq), 7, dtype=np.uint32)
        x = dict(ref_to_query=r2q, is_reverse=bool(i & 1), ref_start=500 * i, ref_end=500 * i + len(sr.refseq), called=called)
        extra.append(x)
        if o.status == api.READ_OK:
            host_in.append((i, dict(x, refseq=sr.refseq, eventAlignment=o.eventAlignment, event_mean=o.event_mean,
                                    shift=o.shift, scale=o.scale, events_per_base=o.eventsPerBase,
                                    event_start=o.event_start, raw=sr.raw)))
    got = b.eventalign_features(extra, window=50, want_records=True)
    want = ctx.eventalign_features([d for _, d in host_in], window=50)
    for i, o in enumerate(got):
        if res[i].status != api.READ_OK:
            assert o["status"] == res[i].status and o["signal"].shape[0] == 0
    for (i, _), w in zip(host_in, want):
        assert got[i]["status"] == api.READ_OK and w["status"] == api.READ_OK
        assert w["signal"].shape[0] > 0
        for key in AP_KEYS + ("event", "ref_pos", "label", "indel"):
            np.testing.assert_array_equal(got[i][key], w[key], err_msg=f"read {i} {key}")
        assert not np.isin(got[i]["coords"], extra[i]["called"]).any()
    b.drop_workspace()
    with pytest.raises(Exception):
        b.eventalign_features(extra)                                  # the resident arrays are gone: call order error
    b.release()


def test_resident_chain_ultra_long_read(ctx, port, pore_mean):
    """BASELINE configs[2] through rows f1/f2: a 130-kb read (~3000 windows in one serial chain, ~1.2*10^5 tensor rows
    written by one CTA in ~1100 chunks) next to a short one, resident chain against the oracle port."""
    ref = synth.make_reference(400_000, 391)
    base = synth.simulate_batch(ref, [130_000, 1500], pore_mean, seed=392)
    b = ctx.upload([api.Read.from_synth(r, use_dac=True) for r in base])
    b.run()
    b.fetch()
    res = b.results()
    extra = []
    for i, sr in enumerate(base):
        q2r = np.asarray(sr.query_to_ref)
        r2q = np.zeros(len(sr.refseq), dtype=np.int32)
        r2q[q2r[q2r >= 0]] = np.nonzero(q2r >= 0)[0]
        extra.append(dict(ref_to_query=r2q, is_reverse=bool(sr.flag & 16), ref_start=sr.pos, ref_end=sr.pos + len(sr.refseq)))
    out = b.eventalign_features(extra, window=50, want_records=True)
    for sr, o, x, f in zip(base, res, extra, out):
        assert o.status == api.READ_OK and f["status"] == api.READ_OK
        rec = port.eventalign(sr.refseq, x["ref_to_query"], o.eventAlignment[:, 0], o.eventAlignment[:, 1],
                              o.event_mean.astype(np.float64), o.shift, o.scale, o.eventsPerBase, pore_mean)
        for key in ("event", "ref_pos", "label", "indel"):
            np.testing.assert_array_equal(f[key], rec[key], err_msg=key)
        want = port.dnn_features(sr.refseq, x["ref_to_query"], x["is_reverse"], x["ref_start"], x["ref_end"], rec,
                                 sr.raw.astype(np.float64), o.event_start, o.shift, o.scale)
        for key in AP_KEYS:
            np.testing.assert_array_equal(f[key], want[key], err_msg=key)
        assert f["signal"].shape[0] > 0.85 * (len(sr.refseq) - 8)
    b.release()


def test_submit_chain_from_concurrent_threads(ctx, ea_golden, golden_reads, golden_v2):
    """dnb_submit_chain (host buffers in, normaliseEvents results + tensors out) called from four host threads at once,
    as the patched OpenMP read loop does: every call returns the reference's golden alignment and tensors."""
    from concurrent.futures import ThreadPoolExecutor
    reads = all_golden_reads(golden_reads, golden_v2)
    # NpzFile loads lazily and is not thread-safe: take what the threads need out of it first
    e = {key: ea_golden[key] for tag, _ in reads for key in
         [f"e_{tag}_strand", f"e_{tag}_ref_to_query"] + [f"e_{tag}_ap_" + k for k in AP_KEYS]}

    def job(k):
        sel = reads[k % 3:] + reads[:k % 3]                                  # a different read order per thread
        ins = [api.Read(None, g.basecall, g.refseq, g.query_to_ref, dac=g.dac, dac_offset=float(synth.DAC_OFFSET),
                        dac_scale=float(synth.DAC_SCALE)) for _, g in sel]
        extra = []
        for tag, g in sel:
            ref_start, ref_end, is_rev = (int(x) for x in e[f"e_{tag}_strand"])
            extra.append(dict(ref_to_query=e[f"e_{tag}_ref_to_query"], is_reverse=is_rev, ref_start=ref_start, ref_end=ref_end))
        b = ctx.submit_chain(ins, extra, window=50)
        res, feats = b.results(), b.feature_results()
        b.release()
        return sel, res, feats

    with ThreadPoolExecutor(max_workers=4) as ex:
        outs = list(ex.map(job, range(8)))
    for sel, res, feats in outs:
        for (tag, g), r, f in zip(sel, res, feats):
            assert r.status == api.READ_OK and f["status"] == api.READ_OK
            np.testing.assert_array_equal(r.eventAlignment, g.align, err_msg=tag)
            for key in AP_KEYS:
                np.testing.assert_array_equal(f[key], e[f"e_{tag}_ap_" + key], err_msg=f"{tag} {key}")
